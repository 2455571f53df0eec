"""Kernel-level breakdown of one engine step (torch profiler over eager steps): GPU microseconds per kernel per step.
    python scripts/profile_step.py [train|vgd|itm|search_weight|search_arch] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmnas_b200
from mmnas_b200 import genotypes
from mmnas_b200.data.synthetic import Cfg, make_batch, init_dict, spec_for, compact
from mmnas_b200.engine import TrainStep, SearchStep, LOSSES, tree_map
from mmnas_b200.model.nets import Net_Full, Net_Search

what = sys.argv[1] if len(sys.argv) > 1 else 'train'
dev = 'cuda'
torch.manual_seed(888)
if what.startswith('search'):
    spec = spec_for('vqa', batch=64)
    cfg = Cfg(mode='search')
    net = Net_Search(cfg, init_dict(spec)).to(dev).train()
    b = tree_map(lambda t: t.to(dev), make_batch(spec, seed=2000))
    step = SearchStep(net, lr_base=cfg.NET_LR_BASE, epoch_steps=10 ** 6)
    fn = (lambda: step.weight_step(*b)) if what == 'search_weight' else (lambda: step.arch_step(*b))
else:
    task = {'train': 'vqa'}.get(what, what)
    arch = {'vqa': 'mmnas_vqa', 'vgd': 'mmnas_vgd', 'itm': 'mmnas_itm'}[task]
    spec = spec_for(task, batch=64)
    cfg = Cfg(genotype=genotypes.shipped(arch), SCORES_LOSS='kld')
    net = Net_Full(cfg, init_dict(spec), task=task).to(dev).train()
    b = tree_map(lambda t: t.to(dev), compact(make_batch(spec, seed=1000)))
    step = TrainStep(net, lr_base=cfg.NET_LR_BASE, epoch_steps=10 ** 6, use_graph=False, loss_fn=LOSSES[task])
    fn = lambda: step(*b)
for _ in range(5): fn()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
N = 5
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(N): fn()
    torch.cuda.synchronize()
skip = ('autograd', 'aten::', 'BackboneFn', 'AttBlock', 'FFNBlock', 'StemImage', 'LayerNormFn', 'LinearFn', 'AddLayerNorm', 'MixedSum', 'Optimizer')
rows = [(e.key, e.device_time_total / N, e.count / N) for e in prof.key_averages() if e.device_time_total > 0 and not e.key.startswith(skip)
        and not e.key.endswith('Backward0') and not e.key.endswith('Backward')]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
ours = sum(r[1] for r in rows if 'anonymous namespace' in r[0] or 'unnamed' in r[0])
print('%s: GPU kernel time %.0f us/step (library kernels %.0f us, other %.0f us), %d launches/step' % (what, tot, ours, tot - ours, sum(r[2] for r in rows)))
for k, us, n in rows[:40]:
    print('%8.1f us %6.1f x %7.1f us  %s' % (us, n, us / n, k[:120]))
if len(sys.argv) > 2:
    json.dump({'workload': what, 'gpu_us_per_step': tot, 'library_us': ours, 'other_us': tot - ours,
               'kernels': [{'kernel': k[:160], 'us_per_step': us, 'launches_per_step': n} for k, us, n in rows]}, open(sys.argv[2], 'w'), indent=1)

if os.environ.get('MMNAS_TIMELINE', '0') == '1':
    # Device timeline: union of kernel intervals (busy), wall span per step, and the largest idle gaps with their neighbours.
    from torch.autograd import DeviceType
    ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == DeviceType.CUDA),
                key=lambda t: t[0])
    span = ev[-1][1] - ev[0][0]
    busy, cur_s, cur_e, gaps = 0.0, ev[0][0], ev[0][1], []
    last_name = ev[0][2]
    for s, e, n in ev[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, last_name[:60], n[:60]))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
        if e >= cur_e: last_name = n
    busy += cur_e - cur_s
    gaps.sort(key=lambda g: -g[0])
    print('timeline: span %.0f us/step, busy (union) %.0f us/step, idle %.0f us/step in %d gaps/step' % (span / N, busy / N, (span - busy) / N, len(gaps) / N))
    hist = {}
    for g, a, b_ in gaps:
        k = (a.split('(')[0][-40:], b_.split('(')[0][-40:])
        hist.setdefault(k, [0, 0.0]); hist[k][0] += 1; hist[k][1] += g
    for k, (c, t) in sorted(hist.items(), key=lambda kv: -kv[1][1])[:25]:
        print('  idle %7.1f us/step over %5.1f gaps/step  after %-40s before %s' % (t / N, c / N, k[0], k[1]))
