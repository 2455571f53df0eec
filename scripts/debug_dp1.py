import sys, collections, torch
sys.path.insert(0, '.')
from tests.test_gpu_dp import _setup
import mmnas_b200
from mmnas_b200 import runtime
from mmnas_b200.engine import FlatGrads, BucketReducer
from mmnas_b200.model.nets import Net_Full
dev = 'cuda'
mmnas_b200.set_precision(sys.argv[1] if len(sys.argv) > 1 else 'fp32'); torch.manual_seed(3)
spec, cfg, init, inputs, target = _setup(8)
net = Net_Full(cfg, init).to(dev).train()
din, dt = tuple(t.to(dev) for t in inputs), target.to(dev)
fg = FlatGrads(net.parameters()); red = BucketReducer(fg, bucket_mb=8.0)
names = {id(p): n for n, p in net.named_parameters()}
red.enabled = True
for i, p in enumerate(fg.params):
    p.register_post_accumulate_grad_hook(red._hooks[i])
events = []
launched = []
red._launch = lambda b: launched.append(b)
counts = collections.Counter()
orig_hooks = list(red._hooks)
def mk(i):
    def h(param):
        counts[names[id(param)]] += 1
        events.append((names[id(param)], red.bucket_of[i]))
        orig_hooks[i](param)
    return h
red._hooks = [mk(i) for i in range(len(fg.params))]
# re-register wrapped hooks for autograd path too
for i, p in enumerate(fg.params):
    p._post_accumulate_grad_hooks.clear() if hasattr(p, '_post_accumulate_grad_hooks') and p._post_accumulate_grad_hooks else None
    p.register_post_accumulate_grad_hook(red._hooks[i])
fg.zero(); red.reset()
runtime.direct_grads, runtime.grad_listener = True, red.notify
loss = torch.nn.functional.binary_cross_entropy_with_logits(net(din), dt, reduction='sum'); loss.backward()
runtime.direct_grads, runtime.grad_listener = False, None
torch.cuda.synchronize()
multi = {k: v for k, v in counts.items() if v != 1}
print('params notified != once:', multi)
missing = [n for n in names.values() if counts[n] == 0]
print('never notified:', missing[:10], len(missing))
print('launch order', launched)
print('bucket 0 params:', [names[id(fg.params[i])] for i in red.buckets[0][2]])
print('first 12 events', events[:12])
