#!/bin/bash
# Round-2 measurement pass on ONE B200 (run through gpurun).  Everything lands in gpurun_out/r02_*; copy what is to be
# judged into profiles/.
set -u
TAG=r02
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
cp gpurun_out/parity.jsonl gpurun_out/${TAG}_parity.jsonl 2>/dev/null
cp gpurun_out/eager_port_timing.json gpurun_out/${TAG}_eager_port_timing.json 2>/dev/null
# ncu traffic of every library kernel over one eager step (Python-composed blocks: one launch per primitive) ...
MMNAS_COMPOSE_PY=1 bash scripts/ncu_traffic.sh > gpurun_out/ncu_traffic.log 2>&1; tail -14 gpurun_out/ncu_traffic.log
cp gpurun_out/ncu_traffic.json gpurun_out/${TAG}_ncu_traffic_primitives.json
cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json       # bench.py reads roofline.traffic from here
# ... the launch list of the default bench command (graph replays included) ...
bash scripts/ncu_launches.sh ${TAG} 4000 1400 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
# ... and one --set full capture of the fused projection + LayerNorm kernel and of the dominant GEMM
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_ln_kernel|gemm_bf16_pair_kernel" -s 30 -c 6 -f -o gpurun_out/${TAG}_full_gemm \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-workloads > /dev/null 2>&1
# bench lines
timeout 600 python bench.py --steps 30 --warmup 5 --profile-out gpurun_out/${TAG}_live_kernel_table.json 2> gpurun_out/bench_n1.err | tail -1 > gpurun_out/${TAG}_bench_n1.json
cut -c1-300 gpurun_out/${TAG}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/bench_ref.err | tail -1 > gpurun_out/${TAG}_bench_reference.json
cut -c1-300 gpurun_out/${TAG}_bench_reference.json
timeout 300 python bench.py --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline --no-workloads 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_n1_fp32.json
cut -c1-200 gpurun_out/${TAG}_bench_n1_fp32.json
for w in train vgd itm search_weight search_arch; do
  timeout 300 python scripts/profile_step.py $w gpurun_out/${TAG}_profile_$w.json 2>&1 | grep -v Warn | head -3 | cut -c1-200
done
timeout 300 python scripts/bench_gemm_ln.py > gpurun_out/${TAG}_gemm_ln.txt 2>&1; cp gpurun_out/gemm_ln_bench.json gpurun_out/${TAG}_gemm_ln.json
timeout 300 python scripts/bench_attention.py > gpurun_out/${TAG}_attention.txt 2>&1; cp gpurun_out/attention_bench.json gpurun_out/${TAG}_attention.json
timeout 300 python scripts/bench_ops.py gpurun_out/${TAG}_ops_roofline.json > gpurun_out/bench_ops.log 2>&1; tail -3 gpurun_out/bench_ops.log | cut -c1-300
timeout 300 python scripts/bench_gemm_pair.py > gpurun_out/${TAG}_gemm_shapes.txt 2>&1
timeout 300 python scripts/bench_gemm.py 2>&1 | grep cublas > gpurun_out/${TAG}_gemm_cublas.txt
timeout 300 python scripts/bench_mining.py gpurun_out/${TAG}_itm_mining.json 2>&1 | tail -3
timeout 300 python scripts/bench_lstm.py > gpurun_out/${TAG}_lstm.txt 2>&1; tail -4 gpurun_out/${TAG}_lstm.txt
timeout 300 python scripts/bench_gemm_bn64.py > gpurun_out/${TAG}_gemm_bn64.txt 2>&1
ls gpurun_out | grep "^${TAG}_" | tr '\n' ' '
