"""Attention core (mmnas_attn_fwd / mmnas_attn_bwd, bf16 arm) at the shapes of the step; CUDA-graph replays."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmnas_b200
from mmnas_b200 import kernels as K
DEV = 'cuda'
res = {}
for (B, h, Nq, Nk, what) in [(64, 8, 100, 100, 'SA_y / RSA_y T'), (64, 8, 100, 14, 'GA_y T'), (64, 8, 14, 14, 'SA_x T'),
                             (64, 4, 100, 100, 'SA_y S'), (192, 8, 36, 36, 'SA_y I'), (192, 8, 36, 50, 'GA_y I'), (192, 8, 50, 50, 'SA_x I')]:
    I = h * 64
    rel = what.startswith('SA_y / RSA') and os.environ.get('ATT_BIAS', '1') == '1'
    sets = []
    for i in range(4):
        qkv = torch.randn(B * max(Nq, Nk), 3 * I, device=DEV).to(torch.bfloat16)
        o = torch.empty(B * Nq, I, device=DEV, dtype=torch.bfloat16)
        do = torch.randn(B * Nq, I, device=DEV).to(torch.bfloat16)
        dqkv = torch.empty_like(qkv)
        bias = torch.randn(B, h, Nq, Nk, device=DEV) if rel else None
        dbias = torch.empty_like(bias) if rel else None
        sets.append((qkv, o, do, dqkv, bias, dbias))
    kmask = torch.zeros(B, Nk, dtype=torch.uint8, device=DEV); kmask[:, Nk - 2:] = 1
    drop = K.Drop(mmnas_b200.runtime.rng_state(DEV), 7, 0.1) if os.environ.get('ATT_P', '0.1') != '0' else K.NO_DROP
    def fwd(s):
        qkv, o, do, dqkv, bias, dbias = s
        K.attn_fwd(B, h, Nq, Nk, qkv[:, 2 * I:].data_ptr(), 3 * I, qkv[:, I:].data_ptr(), 3 * I, qkv.data_ptr(), 3 * I, kmask, bias, o, I, 0.125, drop)
    def bwd(s):
        qkv, o, do, dqkv, bias, dbias = s
        K.attn_bwd(B, h, Nq, Nk, qkv[:, 2 * I:].data_ptr(), 3 * I, qkv[:, I:].data_ptr(), 3 * I, qkv.data_ptr(), 3 * I, kmask, bias, o, I, do, I,
                   dqkv[:, 2 * I:].data_ptr(), 3 * I, dqkv[:, I:].data_ptr(), 3 * I, dqkv.data_ptr(), 3 * I, dbias, 0.125, drop)
    row = {}
    for name, fn in (('fwd', fwd), ('bwd', bwd)):
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
        row[name] = e0.elapsed_time(e1) * 1e3 / 80
    fl = 4.0 * B * h * Nq * Nk * 64
    res['%dx%dx%dx%d %s' % (B, h, Nq, Nk, what)] = row
    print('%-32s fwd %6.2f us (%5.1f TF/s)  bwd %6.2f us (%5.1f TF/s)' % ('%dx%dx%dx%d %s' % (B, h, Nq, Nk, what), row['fwd'], fl / row['fwd'] / 1e6, row['bwd'], 2.5 * fl / row['bwd'] / 1e6))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/attention_bench%s.json' % os.environ.get('TAG', ''), 'w'), indent=1)
