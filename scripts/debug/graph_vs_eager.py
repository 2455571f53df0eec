import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mmnas_b200
from mmnas_b200 import runtime
from mmnas_b200.engine import TrainStep
from mmnas_b200.model.nets import Net_Full
from tests.test_gpu_nets import full_setup
DEV = 'cuda'
torch.manual_seed(1)
spec, cfg, init, inputs, target = full_setup(4, p=0.0)
net0 = Net_Full(cfg, init).to(DEV).train()
din, dt = tuple(t.to(DEV) for t in inputs), target.to(DEV)
runs = {}
with mmnas_b200.precision('bf16'):
    for name, graph in (('eagerA', False), ('eagerB', False), ('graphA', True), ('graphB', True)):
        net = copy.deepcopy(net0)
        step = TrainStep(net, use_graph=graph)
        grads, params, losses = [], [], []
        for k in range(3):
            losses.append(step(din, dt).item())
            torch.cuda.synchronize()
            grads.append(step.grads.flat.clone())
            params.append(torch.cat([p.detach().reshape(-1) for p in net.parameters()]))
        runs[name] = (grads, params, losses, [n for n, _ in net.named_parameters()], step)
def cmp(a, b):
    ga, pa, la, names, sa = runs[a]; gb, pb, lb, _, sb = runs[b]
    print(a, b, 'loss', la, lb)
    for k in range(3):
        dg = (ga[k] - gb[k]).abs().max().item() / ga[k].abs().max().item()
        dp = (pa[k] - pb[k]).abs().max().item()
        print('  step', k, 'grad maxdiff/max %.3e' % dg, 'gradnorm %.4e %.4e' % (ga[k].norm().item(), gb[k].norm().item()), 'param maxabsdiff %.3e' % dp)
    # per-parameter worst at step 0
    fg = sa.grads
    worst = []
    for i, p in enumerate(fg.params):
        o, n = fg.offsets[i], p.numel()
        for k in range(1):
            a_, b_ = ga[k][o:o+n], gb[k][o:o+n]
            worst.append(((a_-b_).abs().max().item() / max(a_.abs().max().item(), 1e-30), names[i], a_.abs().max().item()))
    for w in sorted(worst, reverse=True)[:8]:
        print('   %.3e %s (max %.3e)' % w)
cmp('eagerA', 'eagerB'); cmp('graphA', 'graphB'); cmp('eagerA', 'graphA')
