"""CTA-pair vs single-CTA tcgen05 GEMM on the shapes of the MMnas-VQA train step (batch 64)."""
import os, sys, torch
sys.path.insert(0, '.')
from mmnas_b200 import kernels as K
dev = 'cuda'
def run(M, N, Kd, a_mn, b_mn, out_bf16, split=1, iters=30):
    A = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).bfloat16()
    B = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).bfloat16()
    C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    f = lambda: K.gemm_bf16(M, N, Kd, A, A.stride(0), a_mn, B, B.stride(0), b_mn, C, N, split_k=split)
    for _ in range(3): f()
    torch.cuda.synchronize(); torch.cuda._sleep(int(4e7))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters
print('forward / dgrad shapes (bf16 out)')
for shape in [(6400, 2048, 512), (6400, 1536, 512), (6400, 512, 2048), (6400, 512, 512), (6400, 1024, 512), (896, 512, 512), (896, 1536, 512), (896, 2048, 512), (896, 512, 2048), (896, 1024, 512)]:
    for (a_mn, b_mn) in [(0, 0), (0, 1)]:
        row = []
        for pair in ('0', '1'):
            os.environ['MMNAS_GEMM_PAIR'] = pair
            row.append(run(*shape, a_mn, b_mn, 1))
        print('%s %s%s single %.1f us | pair %.1f us' % (shape, 'mn' if a_mn else 'k', 'mn' if b_mn else 'k', *row))
print('wgrad shapes (mn,mn, fp32 red)')
for (M, N, Kd) in [(512, 512, 6400), (2048, 512, 6400), (512, 2048, 6400), (1536, 512, 6400), (1024, 512, 896), (2048, 512, 896), (1536, 512, 896)]:
    for sk in (1, 2, 3, 4, 6, 9, 12, 18):
        row = []
        for pair in ('0', '1'):
            os.environ['MMNAS_GEMM_PAIR'] = pair
            row.append(run(M, N, Kd, 1, 1, 0, sk))
        print('wgrad %s sk%d single %.1f us | pair %.1f us' % ((M, N, Kd), sk, *row))
