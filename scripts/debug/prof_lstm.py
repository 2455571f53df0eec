import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, mmnas_b200
from mmnas_b200.functional import LSTMFn
from torch.profiler import profile, ProfilerActivity
DEV = 'cuda'
mmnas_b200.set_precision('bf16')
for (B, T, E, H) in [(64, 14, 300, 512), (192, 50, 300, 512)]:
    mod = torch.nn.LSTM(E, H, num_layers=1, batch_first=True).to(DEV)
    emb = torch.randn(B, T, E, device=DEV, requires_grad=True)
    go = torch.randn(B, T, H, device=DEV)
    ps = (mod.weight_ih_l0, mod.weight_hh_l0, mod.bias_ih_l0, mod.bias_hh_l0)
    for _ in range(3):
        LSTMFn.apply(emb, *ps).backward(go)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(5):
            LSTMFn.apply(emb, *ps).backward(go)
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0 and not e.key.startswith(('aten::', 'LSTMFn', 'autograd'))]
    rows.sort(key=lambda r: -r[1])
    print('B=%d T=%d H=%d' % (B, T, H))
    for k, us, n in rows[:14]:
        print('  %8.1f us  x%4.1f  %s' % (us, n, k[:100]))
