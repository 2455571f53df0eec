import cProfile, pstats, sys, io, torch
sys.path.insert(0, '.')
import mmnas_b200
from mmnas_b200.data.synthetic import Cfg, SynthSpec, make_batch, init_dict
from mmnas_b200.engine import SearchStep
from mmnas_b200.model.nets import Net_Search
dev = 'cuda'
mmnas_b200.set_precision('bf16')
torch.manual_seed(888)
spec = SynthSpec(batch=64)
cfg = Cfg(mode='search')
inputs, target = make_batch(spec)
net = Net_Search(cfg, init_dict(spec)).to(dev).train()
step = SearchStep(net)
din, dt = tuple(t.to(dev) for t in inputs), target.to(dev)
for _ in range(3):
    step.weight_step(din, dt)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step.weight_step(din, dt)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[:7000])
