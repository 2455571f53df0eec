import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, mmnas_b200
from mmnas_b200 import genotypes, runtime, functional as Fn
from mmnas_b200.data.synthetic import Cfg, make_batch, init_dict, spec_for
from mmnas_b200.engine import itm_loss
from mmnas_b200.model.nets import Net_Full
sys.path.insert(0, 'tests')
from util import condition_rsa_
DEV = 'cuda'
torch.manual_seed(888)
B = 3
spec = spec_for('itm', batch=B, vocab=1000, n_ans=10)
cfg = Cfg(genotype=genotypes.shipped('mmnas_itm'), DROPOUT_R=0.0)
inputs, _ = make_batch(spec, seed=889)
net = Net_Full(cfg, init_dict(spec), task='itm').train()
with torch.no_grad():
    condition_rsa_(dict(net.named_parameters()))
net = net.to(DEV)
din = tuple(t.to(DEV) for t in inputs)
res = {}
captured = {}
orig_bwd = Fn.LSTMFn.backward
def spy(ctx, dout):
    captured['dout'] = dout.detach().clone()
    x16, wih16, whh16, ws = ctx.saved_tensors
    out = orig_bwd(ctx, dout)
    T, Bq, E, Ep, H = ctx.meta
    TB = T * Bq
    al = lambda v: (v + 255) & ~255
    o_dg = al((TB + Bq) * H * 2) + al(TB * 4 * H * 4) + al(TB * H * 4)
    captured['dg'] = ws[o_dg:o_dg + TB * 4 * H * 2].view(torch.bfloat16).view(TB, 4 * H).float().clone()
    captured['h16'] = ws[:TB * H * 2].view(torch.bfloat16).view(TB, H).float().clone()
    captured['x16'] = x16.float().clone()
    captured['meta'] = ctx.meta
    return out
Fn.LSTMFn.backward = staticmethod(spy)
with mmnas_b200.precision('bf16'):
    for native in (True, False):
        runtime.native_lstm = native
        net.zero_grad()
        loss = itm_loss(net(din))
        loss.backward()
        res[native] = {k: p.grad.detach().clone() for k, p in net.named_parameters() if 'lstm' in k}
nw = lambda a, b: ((a - b).norm() / b.norm()).item()
for k in res[True]:
    print(k, 'native vs cudnn', nw(res[True][k], res[False][k]), 'norm', res[False][k].norm().item())
T, Bq, E, Ep, H = captured['meta']
dg, h16, x16 = captured['dg'], captured['h16'], captured['x16']
ref_hh = dg.t() @ h16
ref_ih = (dg.t() @ x16)[:, :E]
print('our GEMM vs fp32 matmul on the same bf16 operands: hh', nw(res[True]['lstm.weight_hh_l0'], ref_hh), 'ih', nw(res[True]['lstm.weight_ih_l0'], ref_ih))
print('fp32 matmul (bf16 operands) vs cudnn: hh', nw(ref_hh, res[False]['lstm.weight_hh_l0']), 'ih', nw(ref_ih, res[False]['lstm.weight_ih_l0']))
# float64 LSTM reference with the same upstream gradient
ref = torch.nn.LSTM(E, H, num_layers=1, batch_first=True).double().to(DEV)
with torch.no_grad():
    for a, b in zip(ref.parameters(), net.lstm.parameters()): a.copy_(b.double())
emb = net.embedding(din[3]).detach().double().requires_grad_(True)
o, _ = ref(emb)
o.backward(captured['dout'].double())
print('cudnn vs float64 LSTM (same dout as native saw): hh', nw(res[False]['lstm.weight_hh_l0'].double(), ref.weight_hh_l0.grad), 'ih', nw(res[False]['lstm.weight_ih_l0'].double(), ref.weight_ih_l0.grad))
print('native vs float64 LSTM: hh', nw(res[True]['lstm.weight_hh_l0'].double(), ref.weight_hh_l0.grad), 'ih', nw(res[True]['lstm.weight_ih_l0'].double(), ref.weight_ih_l0.grad),
      'b', nw(res[True]['lstm.bias_ih_l0'].double(), ref.bias_ih_l0.grad))
# gate gradients themselves: dG (ours, bf16) against float64 autograd is not directly exposed; check magnitudes over time
dgt = dg.view(T, Bq, 4 * H)
print('|dG_t| per step (first 5, last 5):', [round(float(dgt[t].norm()), 6) for t in list(range(5)) + list(range(T - 5, T))])
print('|dout_t| per step (first 5, last 5):', [round(float(captured['dout'][:, t].norm()), 6) for t in list(range(5)) + list(range(T - 5, T))])
