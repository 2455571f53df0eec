#!/bin/bash
# One tcgen05 GEMM launch under ncu --set full with SASS-level stall sampling (run on the GPU box via gpurun).
set -u
mkdir -p gpurun_out
cat > /tmp/one_gemm.py <<'PY'
import os, sys, torch
sys.path.insert(0, '.')
from mmnas_b200 import kernels as K
M, N, Kd = 6400, 2048, 512
A = torch.randn(M, Kd, device='cuda').bfloat16(); B = torch.randn(N, Kd, device='cuda').bfloat16()
C = torch.zeros(M, N, device='cuda', dtype=torch.bfloat16)
for _ in range(4):
    K.gemm_bf16(M, N, Kd, A, Kd, 0, B, Kd, 0, C, N)
torch.cuda.synchronize()
PY
MMNAS_GEMM_PAIR=${PAIR:-0} MMNAS_GEMM_BN=${BN:-256} timeout 300 ncu --set full --import-source on --clock-control none \
    -k regex:"gemm_bf16" -s 3 -c 1 -o /tmp/one python /tmp/one_gemm.py > /tmp/ncu_one.log 2>&1
tail -3 /tmp/ncu_one.log | cut -c1-200
ncu -i /tmp/one.ncu-rep --page source --csv --print-source sass > gpurun_out/gemm_one_source_${TAG:-a}.csv 2>/tmp/src.err || tail -3 /tmp/src.err
ncu -i /tmp/one.ncu-rep --page details --csv > gpurun_out/gemm_one_details_${TAG:-a}.csv 2>/dev/null
ls -la gpurun_out/gemm_one_* | cut -c1-200
