"""Micro-benchmark of the tcgen05 GEMM: layouts x shapes x epilogues, device time via CUDA events with the GPU parked."""
import os, sys, torch
sys.path.insert(0, '.')
from mmnas_b200 import kernels as K
dev = 'cuda'

def run(M, N, Kd, a_mn, b_mn, out_bf16, accumulate=False, split=1, iters=20):
    A = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).bfloat16()
    B = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).bfloat16()
    C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    f = lambda: K.gemm_bf16(M, N, Kd, A, A.stride(0), a_mn, B, B.stride(0), b_mn, C, N, accumulate=accumulate, split_k=split)
    for _ in range(3): f()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(4e7))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    return us, 2.0 * M * N * Kd / us / 1e6

for bn in ('128', '256'):
    os.environ['MMNAS_GEMM_BN'] = bn
    for (M, N, Kd) in [(6400, 512, 512), (6400, 512, 2048), (6400, 2048, 512), (6400, 1536, 512), (896, 512, 512)]:
        for (a_mn, b_mn) in [(0, 0), (0, 1)]:
            for out_bf16, acc in [(1, False), (0, False), (0, True)]:
                us, tf = run(M, N, Kd, a_mn, b_mn, out_bf16, acc)
                print('BN=%s %5dx%4dx%4d %s%s out=%s acc=%d : %7.1f us %7.1f TF/s' % (bn, M, N, Kd, 'mn' if a_mn else 'k', 'mn' if b_mn else 'k', 'bf16' if out_bf16 else 'f32', acc, us, tf))
os.environ.pop('MMNAS_GEMM_BN')
for (M, N, Kd, sk) in [(512, 512, 6400, 9), (2048, 512, 6400, 2), (512, 2048, 6400, 2), (1536, 512, 6400, 3), (512, 512, 6400, 4), (512,512,6400,18)]:
    us, tf = run(M, N, Kd, 1, 1, 0, False, sk)
    print('wgrad %4dx%4dx%4d sk%d : %7.1f us %7.1f TF/s' % (M, N, Kd, sk, us, tf))

# library reference on the same shapes (torch.matmul -> cuBLASLt), for DESIGN.md's headroom table only
def lib(M, N, Kd, iters=20, wgrad=False):
    if wgrad:
        A = torch.randn(Kd, M, device=dev).bfloat16(); B = torch.randn(Kd, N, device=dev).bfloat16()
        f = lambda: torch.matmul(A.t(), B)
    else:
        A = torch.randn(M, Kd, device=dev).bfloat16(); B = torch.randn(N, Kd, device=dev).bfloat16()
        f = lambda: torch.matmul(A, B.t())
    for _ in range(3): f()
    torch.cuda.synchronize(); torch.cuda._sleep(int(4e7))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    return us, 2.0 * M * N * Kd / us / 1e6
for (M, N, Kd) in [(6400, 512, 512), (6400, 512, 2048), (6400, 2048, 512), (6400, 1536, 512), (896, 512, 512), (896, 1536, 512), (896, 2048, 512)]:
    us, tf = lib(M, N, Kd)
    print('cublas %5dx%4dx%4d : %7.1f us %7.1f TF/s' % (M, N, Kd, us, tf))
for (M, N, Kd) in [(512, 512, 6400), (2048, 512, 6400), (1536, 512, 6400), (512, 2048, 6400)]:
    us, tf = lib(M, N, Kd, wgrad=True)
    print('cublas wgrad %4dx%4dx%4d : %7.1f us %7.1f TF/s' % (M, N, Kd, us, tf))
