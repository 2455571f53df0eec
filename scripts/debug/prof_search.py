import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mmnas_b200
from mmnas_b200.data.synthetic import Cfg, make_batch, init_dict, spec_for
from mmnas_b200.engine import SearchStep, tree_map
from mmnas_b200.model.nets import Net_Search
dev = 'cuda'
torch.manual_seed(888)
spec = spec_for('vqa', batch=64)
cfg = Cfg(mode='search')
net = Net_Search(cfg, init_dict(spec)).to(dev).train()
b = tree_map(lambda t: t.to(dev), make_batch(spec, seed=2000))
step = SearchStep(net, lr_base=cfg.NET_LR_BASE, epoch_steps=10 ** 6)
for _ in range(5): step.weight_step(*b)
torch.cuda.synchronize()
def timeit(fn, n=20):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t) / n * 1e3, (t2 - t) / n * 1e3
print('weight_step host ms, total ms', timeit(lambda: step.weight_step(*b)))
print('arch_step host ms, total ms', timeit(lambda: step.arch_step(*b)))
print('reset_binary_gates', timeit(lambda: net.reset_binary_gates(batched=True)))
from mmnas_b200.model.mixed import MixedOp
MixedOp.MODE = None
net.reset_binary_gates(batched=True); net.unused_modules_off()
with torch.no_grad():
    print('forward only (no_grad)', timeit(lambda: net(b[0])))
print('forward (grad)', timeit(lambda: net(b[0])))
def fb():
    l = step.loss_fn(net(b[0]), b[1]); l.backward()
print('fwd+bwd autograd, no engine', timeit(fb))
net.unused_modules_back()
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step.weight_step(*b)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
