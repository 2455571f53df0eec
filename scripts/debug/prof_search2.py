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
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step.weight_step(*b)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(28)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step.weight_step(*b)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=30, max_name_column_width=60))
print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=25, max_name_column_width=60))
