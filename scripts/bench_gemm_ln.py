"""Fused projection + residual + LayerNorm (mmnas_gemm_ln_bf16) against the two-kernel tail (mmnas_gemm_bf16 +
mmnas_ln_residual_fwd) at the shapes of the step; CUDA-graph replays, buffers rotated through > L2 worth of memory."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmnas_b200
from mmnas_b200 import kernels as K

DEV = 'cuda'
res = {}
for (M, N, Kd, what) in [(6400, 512, 512, 'merge_y T'), (6400, 512, 2048, 'ffn2_y T'), (896, 512, 512, 'merge_x T'),
                         (896, 512, 2048, 'ffn2_x T'), (6400, 256, 256, 'merge_y S'), (6400, 256, 1024, 'ffn2_y S'),
                         (6912, 512, 512, 'merge_y I'), (9600, 512, 2048, 'ffn2_x I')]:
    nset = max(2, int(300e6 // (M * N * 14 + M * Kd * 2)) + 1)
    sets = []
    for i in range(nset):
        A = torch.randn(M, Kd, device=DEV).to(torch.bfloat16)
        x = torch.randn(M, N, device=DEV)
        sets.append((A, x, torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV),
                     torch.empty(M, N, device=DEV, dtype=torch.bfloat16), torch.empty(M, device=DEV), torch.empty(M, device=DEV)))
    W = (torch.randn(N, Kd, device=DEV) / Kd ** 0.5).to(torch.bfloat16)
    bias, gamma, beta = torch.randn(N, device=DEV), torch.ones(N, device=DEV), torch.zeros(N, device=DEV)
    drop = K.Drop(mmnas_b200.runtime.rng_state(DEV), 7, 0.1)

    def fused(s):
        A, x, z, out, out16, mean, sigma = s
        K.gemm_ln_bf16(M, N, Kd, A, Kd, W, Kd, bias, x, gamma, beta, 1e-6, z, out, out16, mean, sigma, drop)

    def unfused(s):
        A, x, z, out, out16, mean, sigma = s
        K.gemm_bf16(M, N, Kd, A, Kd, 0, W, Kd, 0, z, N, bias=bias)
        K.ln_residual_fwd(M, N, x, z, gamma, beta, 1e-6, out, out16, mean, sigma, drop)

    row = {}
    for name, fn in (('fused', fused), ('two_kernels', unfused)):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for s in sets: fn(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for s in sets: fn(s)
            for _ in range(3): g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(20): g.replay()
            e1.record(st)
            torch.cuda.synchronize()
        row[name] = e0.elapsed_time(e1) * 1e3 / (20 * nset)
    row['tflops_fused'] = 2.0 * M * N * Kd / row['fused'] / 1e6
    res['%dx%dx%d %s' % (M, N, Kd, what)] = row
    print('%-28s fused %6.2f us  two kernels %6.2f us  (%.0f TF/s fused)' % ('%dx%dx%d %s' % (M, N, Kd, what), row['fused'], row['two_kernels'], row['tflops_fused']))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/gemm_ln_bench.json', 'w'), indent=1)
