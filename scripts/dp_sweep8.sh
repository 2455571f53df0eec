#!/bin/bash
N=${1:-8}
run() {
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 30 --warmup 5 --no-workloads --no-cpu-baseline > gpurun_out/dp8_$tag.json 2> gpurun_out/dp8_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/dp8_$tag.json"))
    print("%-28s N=%d  %.3f ms/step  %.0f samples/s  e2e %.0f" % ("$tag", d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print("$tag failed", e)
PY
}
run default A=1
run bucket60 MMNAS_BUCKET_MB=60
run bucket110 MMNAS_BUCKET_MB=110
run bucket250 MMNAS_BUCKET_MB=250
run nvls NCCL_ALGO=NVLS
run minctas32 NCCL_MIN_CTAS=32
