"""Data-parallel gradient parity on real GPUs (SURVEY §4 'distributed', §8e): two ranks, each with its own half of a
batch, must end the backward with the same averaged gradients as one process that sees the concatenated batch and
divides its sum-reduced loss by the world size (DDP semantics of train_vqa.py:236-237).  Needs >= 2 GPUs: the
single-GPU round-end run skips it; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu` runs it."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(batch):
    from mmnas_b200 import genotypes
    from mmnas_b200.data.synthetic import Cfg, SynthSpec, make_batch, init_dict
    spec = SynthSpec(batch=batch, vocab=500, n_ans=50)
    cfg = Cfg(genotype=genotypes.shipped('mmnas_vqa'), DROPOUT_R=0.0)
    inputs, target = make_batch(spec, seed=11)
    return spec, cfg, init_dict(spec), inputs, target


def _worker(rank, world, port, q, precision='fp32'):
    import mmnas_b200
    from mmnas_b200 import runtime
    from mmnas_b200.engine import FlatGrads, BucketReducer
    from mmnas_b200.model.nets import Net_Full
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        mmnas_b200.set_precision(precision)
        torch.manual_seed(3)
        spec, cfg, init, inputs, target = _setup(4 * world)
        net = Net_Full(cfg, init).to(dev).train()
        sl = slice(4 * rank, 4 * rank + 4)
        din, dt = tuple(t[sl].to(dev) for t in inputs), target[sl].to(dev)
        fg = FlatGrads(net.parameters())
        red = BucketReducer(fg, bucket_mb=8.0)
        assert red.enabled and len(red.buckets) > 3
        fg.zero()
        red.reset()
        runtime.direct_grads, runtime.grad_listener = True, red.notify
        try:
            loss = torch.nn.functional.binary_cross_entropy_with_logits(net(din), dt, reduction='sum')
            loss.backward()
        finally:
            runtime.direct_grads, runtime.grad_listener = False, None
        red.finish()
        torch.cuda.synchronize()
        q.put((rank, fg.flat.cpu()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.timeout(300)
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_two_rank_gradients_equal_single_process_on_concatenated_batch(precision):
    import mmnas_b200
    from mmnas_b200.engine import FlatGrads
    from mmnas_b200.model.nets import Net_Full
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, precision)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mmnas_b200.set_precision(precision)
    torch.manual_seed(3)
    spec, cfg, init, inputs, target = _setup(4 * world)
    net = Net_Full(cfg, init).to('cuda:0').train()
    fg = FlatGrads(net.parameters())
    fg.zero()
    pred = net(tuple(t.to('cuda:0') for t in inputs))
    (torch.nn.functional.binary_cross_entropy_with_logits(pred, target.to('cuda:0'), reduction='sum') / world).backward()
    ref = fg.flat.cpu()
    mmnas_b200.set_precision('bf16')
    assert torch.equal(got[0], got[1])                                   # every rank holds the same averaged gradients
    if precision == 'fp32':
        scale = ref.abs().max().item()
        assert (got[0] - ref).abs().max().item() < 2e-5 * scale
    else:      # same bf16 operands on both sides; only the fp32 summation order over the batch differs
        assert ((got[0] - ref).norm() / ref.norm()).item() < 2e-3


def _train_worker(rank, world, port, q, use_graph):
    import mmnas_b200
    from mmnas_b200.engine import TrainStep
    from mmnas_b200.model.nets import Net_Full
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    mmnas_b200.set_precision('bf16')
    torch.manual_seed(3)
    spec, cfg, init, inputs, target = _setup(8 * world)
    cfg.DROPOUT_R = 0.1
    net = Net_Full(cfg, init).to(dev).train()
    sl = slice(8 * rank, 8 * rank + 8)
    din, dt = tuple(t[sl].to(dev) for t in inputs), target[sl].to(dev)
    step = TrainStep(net, lr_base=1e-3, epoch_steps=1, bucket_mb=8.0, use_graph=use_graph)
    losses = [float(step(din, dt)) for _ in range(4)]
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().flatten() for p in net.parameters()]).cpu()
    q.put((rank, flat.numpy(), step.grads.flat.cpu().numpy(), losses))    # by value: this process hard-exits
    q.close()
    q.join_thread()                      # flush before the hard exit below
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)          # a captured NCCL graph makes destroy_process_group hang; nothing left to flush


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.timeout(300)
@pytest.mark.parametrize('use_graph', [False, True], ids=['eager', 'cuda_graph'])
def test_two_rank_train_steps_keep_replicas_identical(use_graph):
    """Four bf16 TrainStep iterations (dropout on, side-stream weight gradients, bucketed all-reduce, fused clip+Adam,
    optionally replayed as one CUDA graph): replicas start equal and see different data, so they stay bit-identical
    only if every rank applies the same averaged gradient every step."""
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, q, use_graph)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, flat, grads, losses = q.get(timeout=240)
        got[r] = (torch.from_numpy(flat), torch.from_numpy(grads), losses)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(got[0][1], got[1][1])            # last step's averaged gradients
    assert torch.equal(got[0][0], got[1][0])            # parameters after four updates
    assert got[0][2] != got[1][2]                       # the ranks really saw different batches
    assert all(l == l and l < 1e6 for l in got[0][2] + got[1][2])


def _search_worker(rank, world, port, q):
    import mmnas_b200
    from mmnas_b200.data.synthetic import Cfg, SynthSpec, make_batch, init_dict
    from mmnas_b200.engine import SearchStep
    from mmnas_b200.model.nets import Net_Search
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    mmnas_b200.set_precision('bf16')
    torch.manual_seed(888)
    spec = SynthSpec(batch=8 * world, vocab=500, n_ans=50)
    cfg = Cfg(mode='search', DROPOUT_R=0.1)
    inputs, target = make_batch(spec, seed=11)
    net = Net_Search(cfg, init_dict(spec)).to(dev).train()
    sl = slice(8 * rank, 8 * rank + 8)
    din, dt = tuple(t[sl].to(dev) for t in inputs), target[sl].to(dev)
    torch.manual_seed(100 + rank)            # different generator state per rank: the picks must still agree
    step = SearchStep(net, lr_base=1e-3, epoch_steps=1, bucket_mb=8.0)
    assert step.use_segments and step.reducer.enabled
    losses, picks = [], []
    for _ in range(3):
        losses.append(float(step.weight_step(din, dt)))
        picks.append([m.active_index[0] for m in net.redundant_modules])
        losses.append(float(step.arch_step(din, dt)))
        picks.append([m.active_index[0] for m in net.redundant_modules])
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().flatten() for p in net.parameters()]).cpu()
    q.put((rank, flat.numpy(), picks, losses))
    q.close()
    q.join_thread()
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.timeout(300)
def test_two_rank_search_steps_keep_replicas_identical():
    """Three weight + architecture iterations of SearchStep (segmented replay, sampled path broadcast from rank 0,
    gradient mean over ranks after the stem backward): both replicas walk the same paths and hold bit-identical
    weights and architecture parameters at the end, although they see different data and generator states."""
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_search_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, flat, picks, losses = q.get(timeout=240)
        got[r] = (torch.from_numpy(flat), picks, losses)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1]
    assert len({tuple(p) for p in got[0][1]}) > 1
    assert torch.equal(got[0][0], got[1][0])
    assert got[0][2] != got[1][2]
    assert all(l == l and l < 1e6 for l in got[0][2] + got[1][2])
