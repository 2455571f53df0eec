import os, sys, socket, torch, torch.distributed as dist, torch.multiprocessing as mp
sys.path.insert(0, '.')
from tests.test_gpu_dp import _setup

def worker(rank, world, port, q):
    import mmnas_b200
    from mmnas_b200 import runtime
    from mmnas_b200.engine import FlatGrads, BucketReducer
    from mmnas_b200.model.nets import Net_Full
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    mmnas_b200.set_precision('fp32'); torch.manual_seed(3)
    spec, cfg, init, inputs, target = _setup(4 * world)
    net = Net_Full(cfg, init).to(dev).train()
    sl = slice(4 * rank, 4 * rank + 4)
    din, dt = tuple(t[sl].to(dev) for t in inputs), target[sl].to(dev)
    fg = FlatGrads(net.parameters()); red = BucketReducer(fg, bucket_mb=8.0)
    fired = []
    orig = red._launch
    def launch(b):
        fired.append((b, [red._pending[b]]))
        orig(b)
    red._launch = launch
    fg.zero(); red.reset()
    direct = os.environ.get('DIRECT', '1') == '1'
    runtime.direct_grads, runtime.grad_listener = direct, (red.notify if direct else None)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(net(din), dt, reduction='sum'); loss.backward()
    runtime.direct_grads, runtime.grad_listener = False, None
    n_before = len(fired)
    red.finish(); torch.cuda.synchronize()
    names = [n for n, _ in net.named_parameters()]
    q.put((rank, fg.flat.cpu(), names, fg.offsets, [p.numel() for p in fg.params], n_before, len(red.buckets), [b for b, _ in fired]))
    dist.destroy_process_group()

if __name__ == '__main__':
    world = 2
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn'); q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = dict((r[0], r[1:]) for r in (q.get(timeout=240) for _ in range(world)))
    [p.join() for p in ps]
    f0, names, offs, numels, nb, nbuck, order = res[0]; f1 = res[1][0]
    print('buckets', nbuck, 'launched by hooks', nb, 'order', order[:12])
    bad = 0
    for n, o, k in zip(names, offs, numels):
        d = (f0[o:o + k] - f1[o:o + k]).abs().max().item()
        if d > 0:
            bad += 1
            if bad < 15: print('DIFF', n, d, f0[o:o+k].abs().max().item())
    print('params differing between ranks:', bad, 'of', len(names))
