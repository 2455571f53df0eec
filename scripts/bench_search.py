"""MMnas-VQA supernet search step (BASELINE configs[2]): weight step (sampled path) and arch step (MODE 'full', all
candidates) at B=64, H=256, through engine.SearchStep; device time per step and samples/s."""
import json, sys, torch
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


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


w = timed(lambda: step.weight_step(din, dt))
a = timed(lambda: step.arch_step(din, dt))
mix = (4 * w + (w + a)) / 5          # ALPHA_EVERY = 5: four weight-only iterations, one weight+arch iteration
res = {'weight_step_ms': w, 'arch_step_ms': a, 'weight_samples_s': 64e3 / w, 'arch_samples_s': 64e3 / a,
       'mixed_ms_per_iteration': mix, 'mixed_samples_s': 64e3 / mix, 'note': 'eager launches (the sampled path changes every step)'}
print(json.dumps(res))
json.dump(res, open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/search_step.json', 'w'), indent=1)
