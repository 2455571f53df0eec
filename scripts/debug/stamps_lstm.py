import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, mmnas_b200
from mmnas_b200 import _lib
from mmnas_b200.functional import LSTMFn
mmnas_b200.set_precision('bf16')
lib = _lib.load()
for (B, T, E, H) in [(64, 14, 300, 512), (192, 50, 300, 512)]:
    mod = torch.nn.LSTM(E, H, num_layers=1, batch_first=True).cuda()
    emb = torch.randn(B, T, E, device='cuda', requires_grad=True)
    ps = (mod.weight_ih_l0, mod.weight_hh_l0, mod.bias_ih_l0, mod.bias_hh_l0)
    for mode in ('fwd', 'bwd'):
        for _ in range(2):
            out = LSTMFn.apply(emb, *ps)
            if mode == 'bwd': out.backward(torch.ones_like(out))
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 16)()
        lib.mmnas_debug_lstm_stamps(buf)
        st = list(buf)
        n = 8 if mode == 'fwd' else 10
        print('B=%d T=%d %s' % (B, T, mode), [st[i + 1] - st[i] for i in range(n)], 'total', st[n] - st[0])
