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
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
P = O.leaf_params({k: v.to(DEV) for k, v in net.state_dict().items()}, torch.float64)
inp = tuple((t.double() if t.is_floating_point() else t).to(DEV) for t in inputs)
x, y = O.net_full_vqa(P, inp, cfg.GENOTYPE, return_backbone=True)
x.retain_grad(); y.retain_grad()
xm, ym = O.make_mask(inp[3].unsqueeze(2)), O.make_mask(inp[0])
s = O.head_itm(P, x, y, xm, ym)
a, b, c = s.chunk(3); O.itm_bce_loss(a, b, c).backward()
netd = net.to(DEV)
grabs = {}
orig_head = netd.head
def head(x_out, y_out, x_mask, y_mask):
    x_out.retain_grad(); y_out.retain_grad()
    grabs['x'], grabs['y'], grabs['xm'], grabs['ym'] = x_out, y_out, x_mask, y_mask
    return orig_head(x_out, y_out, x_mask, y_mask)
netd.head = head
with mmnas_b200.precision('fp32'):
    pred = netd(tuple(t.to(DEV) for t in inputs)); loss = itm_loss(pred); loss.backward()
print('fwd x', normwise(grabs['x'], x), 'y', normwise(grabs['y'], y), 'masks eq', torch.equal(grabs['xm'], xm), torch.equal(grabs['ym'], ym))
print('dx', normwise(grabs['x'].grad, x.grad), 'dy', normwise(grabs['y'].grad, y.grad))
dy, dyr = grabs['y'].grad, y.grad
err = (dy.double() - dyr).abs().amax(-1)   # [9,36]
print('dy row err / max', (err / dyr.abs().max()).cpu())
print('ym', ym.view(9, 36).int().sum(1))
for n_ in ['attflat_y.mlp.fc.linear.weight','attflat_y.mlp.fc.linear.bias','attflat_y.mlp.linear.weight','attflat_y.linear_merge.weight','proj.weight','attflat_x.mlp.fc.linear.weight']:
    print(n_, normwise(dict(netd.named_parameters())[n_].grad, P[n_].grad))
# padded rows of y: forward values
yo, yr = grabs['y'].double(), y
print('fwd err per row', ((yo - yr).abs().amax(-1) / yr.abs().max()).cpu())
