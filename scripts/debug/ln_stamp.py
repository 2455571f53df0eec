import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, mmnas_b200
from mmnas_b200 import kernels as K
DEV='cuda'
dbg = torch.zeros(16, dtype=torch.int64, device=DEV)
os.environ['MMNAS_LN_DBG'] = str(dbg.data_ptr())
for (M, N, Kd) in [(896, 512, 512), (6400, 512, 512), (6400, 512, 2048), (6400, 256, 256)]:
    A = torch.randn(M, Kd, device=DEV).to(torch.bfloat16); x = torch.randn(M, N, device=DEV)
    z, out = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    out16 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16); mean, sigma = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    W = (torch.randn(N, Kd, device=DEV) / Kd ** 0.5).to(torch.bfloat16)
    bias, gamma, beta = torch.randn(N, device=DEV), torch.ones(N, device=DEV), torch.zeros(N, device=DEV)
    drop = K.Drop(mmnas_b200.runtime.rng_state(DEV), 7, 0.1)
    for _ in range(3):
        K.gemm_ln_bf16(M, N, Kd, A, Kd, W, Kd, bias, x, gamma, beta, 1e-6, z, out, out16, mean, sigma, drop)
        torch.cuda.synchronize()
    t = dbg.tolist()
    names = ['prologue', 'pdl_wait', 'tfull', 'pass1', 'sync1', 'pass2+sync2', 'waitread', 'pass3', 'out16+wait', 'final sync', 'xwait_begin', 'xwait_end']
    print((M, N, Kd), ' '.join('%s=%.2fus' % (n, v / 1965.0) for n, v in zip(names, t)))
