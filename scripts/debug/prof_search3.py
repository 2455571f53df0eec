import sys, os, time
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
which = sys.argv[1] if len(sys.argv) > 1 else 'weight'
fn = (lambda: step.weight_step(*b)) if which == 'weight' else (lambda: step.arch_step(*b))
for _ in range(5): fn()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20): fn()
torch.cuda.synchronize(); print(which, 'ms/step', (time.perf_counter() - t) / 20 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): fn()
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in ka if e.device_time_total > 0 and not e.key.startswith(('autograd', 'aten::', 'BackboneFn', 'AttBlock', 'FFNBlock', 'StemImage', 'LayerNormFn'))]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print('GPU kernel us/step total %.0f' % tot)
for k, us, n in rows[:32]:
    print('%8.1f us  %5.1f x  %6.1f us each  %s' % (us, n, us / n, k[:110]))
crow = sorted([(e.key, e.self_cpu_time_total / 5, e.count / 5) for e in ka], key=lambda r: -r[1])[:14]
for k, us, n in crow:
    print('CPU %8.1f us  %6.1f x  %s' % (us, n, k[:90]))
