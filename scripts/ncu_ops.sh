#!/bin/bash
# One launch each of the attention / geometry-bias kernels under ncu --set full with SASS-level stall sampling.
# Runs on the GPU box (via gpurun); writes CSV summaries into gpurun_out/ (the .ncu-rep stays in /tmp).
set -u
mkdir -p gpurun_out
cat > /tmp/one_rsa.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import mmnas_b200
from mmnas_b200.model.modules import RelGeometry
from mmnas_b200.utils.ops_adapter import OpsAdapter
class C:
    HSIZE, DROPOUT_R, REL_SIZE = 512, 0.1, 64
mmnas_b200.set_precision('bf16')
dev = 'cuda'
B, Ny = 64, 100
torch.manual_seed(0)
y = torch.randn(B, Ny, 512, device=dev, requires_grad=True)
ym = torch.zeros(B, 1, 1, Ny, dtype=torch.bool, device=dev); ym[:, :, :, 80:] = True
g4 = torch.randn(B, Ny, Ny, 4, device=dev)
lin = torch.nn.Linear(4, 64).to(dev)
op = OpsAdapter().OPS['rel_self_att_64'](C(), True, True).to(dev).train()
geo = RelGeometry(g4, lin)
for _ in range(3):
    out = op(y, None, ym, None, geo)
    out.sum().backward()
torch.cuda.synchronize()
PY
for K in ${KERNELS:-relbias_fwd relbias_bwd attn_fwd_tc attn_bwd_tc}; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:"$K" -s ${SKIP:-2} -c 1 -o /tmp/op_$K -f python /tmp/one_rsa.py > /tmp/ncu_$K.log 2>&1
  tail -1 /tmp/ncu_$K.log | cut -c1-150
  ncu -i /tmp/op_$K.ncu-rep --page source --csv --print-source sass > gpurun_out/src_$K${SUFFIX:-}.csv 2>/dev/null
  ncu -i /tmp/op_$K.ncu-rep --page details --csv > gpurun_out/det_$K${SUFFIX:-}.csv 2>/dev/null
done
ls -la gpurun_out/src_* gpurun_out/det_* | cut -c20-200
