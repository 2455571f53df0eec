#!/bin/bash
# Data-parallel fixed-cost sweep at N GPUs: NCCL CTA footprint and gradient bucket size (device-resident `value`).
N=${1:-2}
run() {
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 --no-workloads --no-cpu-baseline > gpurun_out/dp_sweep_$tag.json 2> gpurun_out/dp_sweep_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/dp_sweep_$tag.json"))
    print("%-28s N=%d  %.3f ms/step  %.0f samples/s  e2e %.0f" % ("$tag", d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print("$tag failed", e)
PY
}
run default A=1
run maxctas8 NCCL_MAX_CTAS=8
run maxctas4 NCCL_MAX_CTAS=4
run maxctas2 NCCL_MAX_CTAS=2
run bucket60 MMNAS_BUCKET_MB=60
run bucket250 MMNAS_BUCKET_MB=250
run bucket60_ctas4 MMNAS_BUCKET_MB=60 NCCL_MAX_CTAS=4
