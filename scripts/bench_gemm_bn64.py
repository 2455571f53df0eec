"""64-wide N tiles on the text-stream GEMM shapes: correctness against torch.matmul and device time per launch for
BN = 64 / 128 (MMNAS_GEMM_BN forces the tile), every operand layout the step uses."""
import os, sys, torch
sys.path.insert(0, '.')
from mmnas_b200 import kernels as K
dev = 'cuda'
torch.manual_seed(0)
def run(M, N, Kd, a_mn, b_mn, out_bf16, split=1, iters=30):
    A = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).bfloat16()
    B = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).bfloat16()
    C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    f = lambda: K.gemm_bf16(M, N, Kd, A, A.stride(0), a_mn, B, B.stride(0), b_mn, C, N, split_k=split)
    f(); torch.cuda.synchronize()
    ref = (A.float().t() if a_mn else A.float()) @ (B.float() if b_mn else B.float().t())
    err = ((C.float() - ref).norm() / ref.norm()).item()
    for _ in range(3): f()
    torch.cuda.synchronize(); torch.cuda._sleep(int(4e7))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        if split > 1: pass
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters, err
os.environ['MMNAS_GEMM_PAIR'] = '0'
for shape in [(896, 512, 512), (896, 1536, 512), (896, 2048, 512), (896, 512, 2048), (896, 1024, 512), (896, 512, 1024), (896, 512, 1536), (64, 1024, 512), (64, 3136, 1024), (3200, 512, 512)]:
    for (a_mn, b_mn) in [(0, 0), (0, 1)]:
        row = []
        for bn in ('128', '64'):
            os.environ['MMNAS_GEMM_BN'] = bn
            if shape[1] % 64: row += [float('nan'), float('nan')]; continue
            row += list(run(*shape, a_mn, b_mn, 1))
        print('%s %s%s  BN128 %.1f us (err %.1e) | BN64 %.1f us (err %.1e)' % (shape, 'mn' if a_mn else 'k', 'mn' if b_mn else 'k', *row))
for (M, N, Kd) in [(1024, 512, 896), (2048, 512, 896), (512, 2048, 896), (1536, 512, 896), (512, 512, 896)]:
    for sk in (1, 2, 3):
        row = []
        for bn in ('128', '64'):
            os.environ['MMNAS_GEMM_BN'] = bn
            row.append(run(M, N, Kd, 1, 1, 0, sk)[0])
        print('wgrad %s sk%d BN128 %.1f us | BN64 %.1f us' % ((M, N, Kd), sk, *row))
