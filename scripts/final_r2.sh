#!/bin/bash
# Final numbers of the round on ONE B200: full GPU test suite, smoke, the default bench line and the reference arm.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest.log 2>&1; tail -2 gpurun_out/r02_pytest.log
cp gpurun_out/parity.jsonl gpurun_out/r02_parity.jsonl
cp gpurun_out/eager_port_timing.json gpurun_out/r02_eager_port_timing.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 30 --warmup 5 --profile-out gpurun_out/r02_live_kernel_table.json 2> gpurun_out/bench_n1.err | tail -1 > gpurun_out/r02_bench_n1.json
cut -c1-200 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/bench_ref.err | tail -1 > gpurun_out/r02_bench_reference.json
cut -c1-200 gpurun_out/r02_bench_reference.json
timeout 300 python scripts/bench_attention.py > gpurun_out/r02_attention.txt 2>&1; cp gpurun_out/attention_bench.json gpurun_out/r02_attention.json
timeout 300 python scripts/profile_step.py train gpurun_out/r02_profile_train.json 2>&1 | grep -v Warn | head -2 | cut -c1-200
timeout 300 python scripts/profile_step.py itm gpurun_out/r02_profile_itm.json 2>&1 | grep -v Warn | head -2 | cut -c1-200
timeout 300 python scripts/profile_step.py vgd gpurun_out/r02_profile_vgd.json 2>&1 | grep -v Warn | head -2 | cut -c1-200
