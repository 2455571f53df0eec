run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 60 --warmup 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', round(d['value']), round(d['ms_per_step'],3))"; }
run 29551 default
MMNAS_BUCKET_MB=8 run 29552 bucket8
MMNAS_BUCKET_MB=64 run 29553 bucket64
NCCL_MAX_CTAS=8 run 29554 maxctas8
NCCL_MAX_CTAS=4 MMNAS_BUCKET_MB=64 run 29555 maxctas4_b64
