import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, mmnas_b200
from mmnas_b200 import kernels as K
DEV='cuda'
for (M, N, Kd) in [(896, 512, 512), (6400, 512, 512), (6400, 512, 2048)]:
    A = torch.randn(M, Kd, device=DEV).to(torch.bfloat16); x = torch.randn(M, N, device=DEV)
    z, out = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    out16 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16); mean, sigma = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    W = (torch.randn(N, Kd, device=DEV) / Kd ** 0.5).to(torch.bfloat16)
    bias, gamma, beta = torch.randn(N, device=DEV), torch.ones(N, device=DEV), torch.zeros(N, device=DEV)
    drop = K.Drop(mmnas_b200.runtime.rng_state(DEV), 7, 0.1)
    for _ in range(3):
        K.gemm_ln_bf16(M, N, Kd, A, Kd, W, Kd, bias, x, gamma, beta, 1e-6, z, out, out16, mean, sigma, drop)
        torch.cuda.synchronize()
        K.gemm_bf16(M, N, Kd, A, Kd, 0, W, Kd, 0, z, N, bias=bias)
        K.ln_residual_fwd(M, N, x, z, gamma, beta, 1e-6, out, out16, mean, sigma, drop)
        torch.cuda.synchronize()
