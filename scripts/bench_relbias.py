"""RSA geometry-bias kernels at the MMnas-VQA shape (B=64, N=100, 8 heads): mode 0 (fp32 FFMA2) vs mode 1 (mma.sync),
forward and backward, device us per call replayed from a CUDA graph."""
import sys, torch
sys.path.insert(0, '.')
from mmnas_b200 import kernels as K
dev = 'cuda'
B, N, h, R = 64, 100, 8, 64
torch.manual_seed(0)
g4 = torch.randn(B, N, N, 4, device=dev)
Wy, by = 0.5 * torch.randn(R, 4, device=dev), 0.1 * torch.randn(R, device=dev)
Wr, br = 0.03 * torch.randn(h, R, device=dev), torch.ones(h, device=dev)
bias = torch.empty(B, h, N, N, device=dev)
go = torch.randn(B, h, N, N, device=dev)
dWr, dbr, dWy, dby = torch.zeros_like(Wr), torch.zeros_like(br), torch.zeros_like(Wy), torch.zeros_like(by)


def timed(fn, iters=20, reps=4):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * reps)


for mode in (0, 1):
    f = timed(lambda: K.relbias_fwd(B, N, h, R, None, g4, Wy, by, Wr, br, bias, mode=mode))
    b = timed(lambda: K.relbias_bwd(B, N, h, R, None, g4, Wy, by, Wr, br, go, None, dWy, dby, dWr, dbr, mode=mode))
    print('mode %d: fwd %.1f us, bwd %.1f us' % (mode, f, b))
