#!/bin/bash
# Round-end measurement pass on ONE B200 (run through gpurun): ncu launch list / traffic / full metrics first (the bench
# reads profiles/ncu_traffic.json for roofline.traffic), then the bench lines, per-op roofline, search step, GEMM tables.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
bash scripts/ncu_traffic.sh > gpurun_out/ncu_traffic.log 2>&1; tail -12 gpurun_out/ncu_traffic.log
cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json
bash scripts/ncu_capture.sh $TAG > gpurun_out/ncu_capture.log 2>&1; tail -3 gpurun_out/ncu_capture.log
timeout 400 python bench.py --profile-out gpurun_out/${TAG}_live_kernel_table.json 2> gpurun_out/bench_n1.err | tail -1 > gpurun_out/${TAG}_bench_n1.json
cut -c1-260 gpurun_out/${TAG}_bench_n1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_ref.err | tail -1 > gpurun_out/${TAG}_bench_reference.json
cut -c1-260 gpurun_out/${TAG}_bench_reference.json
timeout 300 python bench.py --precision fp32 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_n1_fp32.json
cut -c1-200 gpurun_out/${TAG}_bench_n1_fp32.json
timeout 300 python scripts/bench_ops.py gpurun_out/${TAG}_ops_roofline.json > gpurun_out/bench_ops.log 2>&1; tail -3 gpurun_out/bench_ops.log | cut -c1-300
timeout 300 python scripts/bench_search.py gpurun_out/${TAG}_search_step.json 2>&1 | tail -1 | cut -c1-400
timeout 300 python scripts/bench_gemm_pair.py > gpurun_out/${TAG}_gemm_shapes.txt 2>&1
timeout 300 python scripts/bench_gemm.py 2>&1 | grep cublas > gpurun_out/${TAG}_gemm_cublas.txt
timeout 120 python scripts/bench_gemm_dbg.py > gpurun_out/${TAG}_gemm_breakdown.txt 2>&1
cp gpurun_out/parity.jsonl gpurun_out/${TAG}_parity.jsonl 2>/dev/null
ls gpurun_out | grep "^${TAG}_" | tr '\n' ' '
