"""Where does the tcgen05 GEMM spend its time?  MMNAS_GEMM_DEBUG=2 drops the MMAs, =3 drops the epilogue work
(the accumulator is released unread): full vs no-MMA vs no-epilogue device time per shape."""
import os, sys, torch
sys.path.insert(0, '.')
from mmnas_b200 import kernels as K
dev = 'cuda'
def run(M, N, Kd, out_bf16, iters=30):
    A = torch.randn(M, Kd, device=dev).bfloat16(); B = torch.randn(N, Kd, device=dev).bfloat16()
    C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    f = lambda: K.gemm_bf16(M, N, Kd, A, Kd, 0, B, Kd, 0, C, N)
    for _ in range(3): f()
    torch.cuda.synchronize(); torch.cuda._sleep(int(4e7))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters
for bn in ('128', '256'):
    os.environ['MMNAS_GEMM_BN'] = bn; os.environ['MMNAS_GEMM_PAIR'] = '0'
    for shape in [(6400, 2048, 512), (6400, 1536, 512), (6400, 512, 2048), (6400, 512, 512), (896, 512, 512), (896, 2048, 512)]:
        row = []
        for dbg in ('0', '2', '3'):
            os.environ['MMNAS_GEMM_DEBUG'] = dbg
            row.append(run(*shape, 1))
        print('BN=%s %s full %.1f us | no-mma %.1f | no-epilogue %.1f' % (bn, shape, *row))

# fixed cost: one tile, one k-block
os.environ['MMNAS_GEMM_DEBUG'] = '0'
for shape in [(128, 128, 64), (128, 128, 512), (896, 128, 64), (6400, 128, 64), (6400, 512, 64)]:
    print('fixed-cost probe', shape, '%.1f us' % run(*shape, 1))
# an empty kernel through the same launch path
x = torch.zeros(4, device=dev)
st = torch.zeros(2, dtype=torch.int64, device=dev)
def empty():
    K.rng_advance(st)
for _ in range(3): empty()
torch.cuda.synchronize(); torch.cuda._sleep(int(4e7))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): empty()
e1.record(); torch.cuda.synchronize()
print('1-thread kernel back to back: %.2f us' % (e0.elapsed_time(e1) * 1e3 / 50))
