"""Per-op roofline table (BASELINE.md §3 / SURVEY §8d): each fused block at config T (B=64, H=512, 8 heads, Ny=100, Nx=14),
forward and forward+backward device time through the public op classes, against the ALGORITHMIC FLOPs / bytes
(padding and recompute excluded) and the measured peaks.  Also the HBM-bound pieces at config S (H=256).
Every timing replays a CUDA graph of the call (eager calls of these blocks are host-bound: ~10 us of Python per launch);
the HBM-bound pieces rotate over buffer sets larger than the 126 MB L2.
Writes JSON to the path given as argv[1] (default gpurun_out/ops_roofline.json)."""
import json, os, sys, torch
sys.path.insert(0, '.')
import mmnas_b200
from mmnas_b200 import kernels as K
from mmnas_b200.model.modules import RelGeometry
from mmnas_b200.utils.ops_adapter import OpsAdapter

dev = 'cuda'
PEAK = json.load(open('MEASURED_PEAKS.json')) if os.path.exists('MEASURED_PEAKS.json') else {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}


class C:
    def __init__(self, h):
        self.HSIZE, self.DROPOUT_R, self.REL_SIZE = h, 0.1, 64


def timed(fn, iters=20, reps=1):
    """device us per call of fn(i) (i = call index, for buffer rotation), replayed from a CUDA graph holding `reps` calls"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(max(3, reps)):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * reps)      # us


def gf_block(name, B, Nq, Nk, H):
    """algorithmic forward GFLOP (2MNK per GEMM), SURVEY §8a"""
    Mq, Mk, h = B * Nq, B * Nk, H // 64
    if name == 'feed_forward':
        return 2 * 2 * Mq * H * 4 * H / 1e9
    proj = 2 * Mq * H * H * 2 + 2 * Mk * H * H * 2          # q + merge on queries, k + v on keys
    att = 4 * B * h * Nq * Nk * 64
    rel = 2 * B * Nq * Nq * 64 * h if name == 'rel_self_att_64' else 0
    return (proj + att + rel) / 1e9


rows = []
mmnas_b200.set_precision('bf16')
B, H, Ny, Nx = 64, 512, 100, 14
torch.manual_seed(0)
y = torch.randn(B, Ny, H, device=dev, requires_grad=True)
x = torch.randn(B, Nx, H, device=dev, requires_grad=True)
ym = torch.zeros(B, 1, 1, Ny, dtype=torch.bool, device=dev); ym[:, :, :, 80:] = True
xm = torch.zeros(B, 1, 1, Nx, dtype=torch.bool, device=dev); xm[:, :, :, 10:] = True
g4 = torch.randn(B, Ny, Ny, 4, device=dev)
lin = torch.nn.Linear(4, 64).to(dev)
cases = [('SA_y', 'self_att_64', y, None, ym, None, Ny, Ny), ('RSA_y', 'rel_self_att_64', y, None, ym, None, Ny, Ny),
         ('GA_y', 'guided_att_64', y, x, ym, xm, Ny, Nx), ('FFN_y', 'feed_forward', y, None, ym, None, Ny, Ny),
         ('SA_x', 'self_att_64', x, None, xm, None, Nx, Nx), ('FFN_x', 'feed_forward', x, None, xm, None, Nx, Nx)]
for label, name, s, pre, sm, pm, nq, nk in cases:
    op = OpsAdapter().OPS[name](C(H), True, True).to(dev).train()
    rel = RelGeometry(g4, lin) if name == 'rel_self_att_64' else None
    go = torch.randn(B, nq, H, device=dev)

    def fwd(i):
        with torch.no_grad():
            op(s, pre, sm, pm, rel)

    def fwdbwd(i):
        out = op(s, pre, sm, pm, rel)
        out.backward(go)

    t_f, t_fb = timed(fwd), timed(fwdbwd)
    gf = gf_block(name, B, nq, nk, H)
    rows.append({'op': label, 'fwd_us': t_f, 'fwd_bwd_us': t_fb, 'fwd_gflop': gf, 'fwd_tflops': gf / t_f * 1e3,
                 'fwd_frac_of_bf16_peak': gf / t_f * 1e3 / PEAK['bf16_tflops'],
                 'fwd_bwd_tflops': 3 * gf / t_fb * 1e3, 'fwd_bwd_frac_of_bf16_peak': 3 * gf / t_fb * 1e3 / PEAK['bf16_tflops']})

# HBM-bound pieces: NSET independent buffer sets (> L2 in total) visited round-robin inside one graph
M, Hs, NSET = 6400, 512, 8
sets = []
for _ in range(NSET):
    sets.append(dict(xx=torch.randn(M, Hs, device=dev), br=torch.randn(M, Hs, device=dev), out=torch.empty(M, Hs, device=dev),
                     out16=torch.empty(M, Hs, device=dev, dtype=torch.bfloat16), mean=torch.empty(M, device=dev),
                     sig=torch.empty(M, device=dev), dz=torch.empty(M, Hs, device=dev),
                     db16=torch.empty(M, Hs, device=dev, dtype=torch.bfloat16)))
a2, b2 = torch.ones(Hs, device=dev), torch.zeros(Hs, device=dev)
da, dbb = torch.zeros(Hs, device=dev), torch.zeros(Hs, device=dev)


def ln_f(i):
    d = sets[i % NSET]
    K.ln_residual_fwd(M, Hs, d['xx'], d['br'], a2, b2, 1e-6, d['out'], d['out16'], d['mean'], d['sig'])


def ln_b(i):
    d = sets[i % NSET]
    K.ln_residual_bwd(M, Hs, d['out'], d['br'], d['mean'], d['sig'], a2, 1e-6, d['dz'], d['db16'], da, dbb)


t = timed(ln_f, reps=NSET)
byts = M * Hs * (4 * 4 + 2)
rows.append({'op': 'residual+LN fwd (y, bf16 arm)', 'us': t, 'bytes': byts, 'gbs': byts / t / 1e3, 'frac_of_hbm_peak': byts / t / 1e3 / PEAK['hbm_gbs']})
t = timed(ln_b, reps=NSET)
byts = M * Hs * (4 * 3 + 2)
rows.append({'op': 'residual+LN bwd (y, bf16 arm)', 'us': t, 'bytes': byts, 'gbs': byts / t / 1e3, 'frac_of_hbm_peak': byts / t / 1e3 / PEAK['hbm_gbs']})
n = 64 * 100 * 256
msets = [dict(outs=[torch.randn(n, device=dev) for _ in range(4)], o=torch.empty(n, device=dev), d0=torch.empty(n, device=dev))
         for _ in range(6)]
gate, gg = torch.tensor([0., 1., 0., 0.], device=dev), torch.empty(4, device=dev)
t = timed(lambda i: K.mixed_accum(msets[i % 6]['outs'], gate, msets[i % 6]['o']), reps=6)
rows.append({'op': 'mixed-op accumulate (S, K=4)', 'us': t, 'bytes': 5 * n * 4, 'gbs': 5 * n * 4 / t / 1e3, 'frac_of_hbm_peak': 5 * n * 4 / t / 1e3 / PEAK['hbm_gbs']})
t = timed(lambda i: K.mixed_alpha_dot(msets[i % 6]['outs'], gate, msets[i % 6]['o'], gg, [None, msets[i % 6]['d0'], None, None]), reps=6)
rows.append({'op': 'mixed-op alpha-dot (S, K=4)', 'us': t, 'bytes': 6 * n * 4, 'gbs': 6 * n * 4 / t / 1e3, 'frac_of_hbm_peak': 6 * n * 4 / t / 1e3 / PEAK['hbm_gbs']})
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/ops_roofline.json'
json.dump({'peaks': PEAK, 'rows': rows}, open(path, 'w'), indent=1)
for r in rows:
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})
