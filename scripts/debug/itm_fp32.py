import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mmnas_b200
from oracle import mmnas_oracle as O
from tests.util import normwise, condition_rsa_
from mmnas_b200 import genotypes
from mmnas_b200.data.synthetic import Cfg, make_batch, init_dict, spec_for
from mmnas_b200.engine import itm_loss
from mmnas_b200.model.nets import Net_Full
DEV='cuda'
torch.manual_seed(888)
B = 3
spec = spec_for('itm', batch=B, vocab=1000, n_ans=10)
cfg = Cfg(genotype=genotypes.shipped('mmnas_itm'), DROPOUT_R=0.0)
inputs, _ = make_batch(spec)
net = Net_Full(cfg, init_dict(spec), task='itm').train()
with torch.no_grad():
    condition_rsa_(dict(net.named_parameters()))
def run_oracle(dtype, dev, stacked):
    P = O.leaf_params({k: v.to(dev) for k, v in net.state_dict().items()}, dtype)
    inp = tuple((t.to(dtype) if t.is_floating_point() else t).to(dev) for t in inputs)
    if stacked:
        s = O.net_full(P, inp, cfg.GENOTYPE, task='itm')
        a, b, c = s.chunk(3)
        loss = O.itm_bce_loss(a, b, c); loss.backward(); sc = s.detach()
    else:
        thirds = [tuple(t[k * B:(k + 1) * B] for t in inp) for k in range(3)]
        loss, sc = O.train_step_itm(P, thirds[0], thirds[1], thirds[2], cfg.GENOTYPE); sc = torch.cat(sc)
    return P, sc, loss
P64, s64, l64 = run_oracle(torch.float64, 'cpu', False)
P64s, s64s, l64s = run_oracle(torch.float64, 'cpu', True)
P32, s32, l32 = run_oracle(torch.float32, 'cpu', False)
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
P32g, s32g, l32g = run_oracle(torch.float32, 'cuda', True)
netd = net.to(DEV)
with mmnas_b200.precision('fp32'):
    pred = netd(tuple(t.to(DEV) for t in inputs)); loss = itm_loss(pred); loss.backward()
print('scores', s64.tolist()); print('ours', pred.tolist())
gmax = max(p.grad.abs().max().item() for p in P64.values() if p.grad is not None)
rows = []
for n_, p_ in netd.named_parameters():
    if P64[n_].grad is None: continue
    r = P64[n_].grad
    rows.append((n_, normwise(p_.grad, r, 1e-2*gmax), normwise(P64s[n_].grad, r, 1e-2*gmax), normwise(P32[n_].grad, r, 1e-2*gmax), normwise(P32g[n_].grad.cpu(), r, 1e-2*gmax)))
print('%-70s %10s %10s %10s %10s' % ('param', 'ours', 'or64stack', 'or32cpu', 'or32gpu'))
for r in rows:
    if max(r[1:]) > 2e-5: print('%-70s %10.2e %10.2e %10.2e %10.2e' % r)
print('n', len(rows), 'loss', float(l64), float(l32), float(l32g), float(loss))
