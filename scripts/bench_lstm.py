"""Text-stem LSTM: functional.LSTMFn (persistent kernels) against torch.nn.LSTM (cuDNN) on the same GPU, forward and
forward+backward, CUDA-graph replays; VQA (B=64, T=14) and ITM (B=192, T=50) shapes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmnas_b200
from mmnas_b200.functional import LSTMFn
DEV = 'cuda'
mmnas_b200.set_precision('bf16')
res = {}
for (B, T, E, H, what) in [(64, 14, 300, 512, 'VQA'), (64, 15, 300, 512, 'VGD'), (192, 50, 300, 512, 'ITM'), (64, 14, 300, 256, 'search')]:
    torch.manual_seed(0)
    mod = torch.nn.LSTM(E, H, num_layers=1, batch_first=True).to(DEV)
    emb = torch.randn(B, T, E, device=DEV, requires_grad=True)
    go = torch.randn(B, T, H, device=DEV)
    ps = (mod.weight_ih_l0, mod.weight_hh_l0, mod.bias_ih_l0, mod.bias_hh_l0)
    def ours(bwd):
        out = LSTMFn.apply(emb, *ps)
        if bwd: out.backward(go)
    def cudnn(bwd):
        out, _ = mod(emb)
        if bwd: out.backward(go)
    row = {}
    for name, fn in (('ours', ours), ('cudnn', cudnn)):
        for bwd in (False, True):
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(3):
                    fn(bwd)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    fn(bwd)
                for _ in range(3): g.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(20): g.replay()
                e1.record(st)
                torch.cuda.synchronize()
            row['%s_%s_us' % (name, 'fwd_bwd' if bwd else 'fwd')] = e0.elapsed_time(e1) * 1e3 / 20
    res['%s B=%d T=%d H=%d' % (what, B, T, H)] = row
    print('%-24s ours fwd %7.1f us  fwd+bwd %7.1f us | cuDNN fwd %7.1f us  fwd+bwd %7.1f us' % (
        '%s B=%d T=%d H=%d' % (what, B, T, H), row['ours_fwd_us'], row['ours_fwd_bwd_us'], row['cudnn_fwd_us'], row['cudnn_fwd_bwd_us']))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/r02_lstm.json', 'w'), indent=1)
