"""Host-side data-parallel logic (FlatGrads + BucketReducer + WarmupAdam) on CPU with the gloo backend, world size 2:
averaged gradients must equal the single-process gradients of the concatenated batch (sum-reduced loss / world),
which is DDP's semantics in train_vqa.py:236."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmnas_b200.engine import FlatGrads, BucketReducer, WarmupAdam


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(),
                               torch.nn.Linear(32, 4), torch.nn.Linear(4, 4))   # last layer stays unused below


def _forward(m, x):
    return m[4](m[3](m[2](m[1](m[0](x)))))      # m[5] never runs: its grads must still arrive as zeros


class _DirectLinear(torch.autograd.Function):
    """What the block backwards do in direct-gradient mode: accumulate into p.grad in place, tell the reducer through
    runtime.notify_grads and hand autograd None (autograd STILL fires the post-accumulate hook for that None)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.set_materialize_grads(False)
        ctx.params = (w, b)
        ctx.save_for_backward(x)
        return x @ w.detach().t() + b.detach()

    @staticmethod
    def backward(ctx, g):
        from mmnas_b200 import runtime
        x, = ctx.saved_tensors
        w, b = ctx.params
        w.grad += g.t() @ x
        b.grad += g.sum(0)
        runtime.notify_grads([w, b])
        return g @ w.detach(), None, None


def _forward_direct(m, x):
    for i in (0, 2, 4):
        x = _DirectLinear.apply(x, m[i].weight, m[i].bias)
        if i < 4:
            x = torch.relu(x)
    return x


def _worker(rank, world, port, q, direct=False):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        m = _model()
        fg = FlatGrads(m.parameters())
        red = BucketReducer(fg, bucket_mb=0.0004)          # tiny buckets -> several all-reduces
        assert len(red.buckets) > 2
        g = torch.Generator().manual_seed(1)
        x = torch.randn(8, 16, generator=g)
        y = torch.randn(8, 4, generator=g)
        xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
        for _ in range(2):                                  # second pass checks re-arming
            fg.zero()
            red.reset()
            if direct:
                from mmnas_b200 import runtime
                runtime.grad_listener = red.notify
                xs = xs.clone().requires_grad_(True)
                loss = ((_forward_direct(m, xs) - ys) ** 2).sum()
                loss.backward()
                runtime.grad_listener = None
            else:
                loss = ((_forward(m, xs) - ys) ** 2).sum()
                loss.backward()
            early = [b for b, (_, _, ids) in enumerate(red.buckets) if red._launched[b] and red._pending[b] != 0]
            assert not early, f'buckets {early} were reduced before all their parameters reported'
            red.finish()
        q.put((rank, fg.flat.clone()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize('direct', [False, True], ids=['autograd_grads', 'direct_grads'])
def test_bucket_reducer_matches_single_process_gradients(direct):
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, direct)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    m = _model()
    fg = FlatGrads(m.parameters())
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 16, generator=g), torch.randn(8, 4, generator=g)
    fg.zero()
    (((_forward(m, x) - y) ** 2).sum() / world).backward()
    assert torch.allclose(got[0], got[1])
    assert torch.allclose(got[0], fg.flat, rtol=1e-5, atol=1e-6)
    unused = fg.view(len(fg.params) - 1)
    assert float(unused.abs().max()) == 0.0                 # the unused layer received a (zero) gradient


def test_flat_grads_survive_grad_none_and_warmup_schedule():
    m = _model()
    fg = FlatGrads(m.parameters())
    for p in m.parameters():
        p.grad = None                                       # what MixedOp.binarize() does to candidate params
    fg.zero()
    assert all(p.grad is not None and p.grad.data_ptr() == fg.view(i).data_ptr() for i, p in enumerate(fg.params))
    opt = WarmupAdam(m.parameters(), lr_base=1.0, epoch_steps=10)
    rates = [opt.rate(s) for s in (1, 10, 11, 20, 21, 30, 31, 1000)]
    assert rates == [0.25, 0.25, 0.5, 0.5, 0.75, 0.75, 1.0, 1.0]      # optimizer.py:25-44
    opt.decay(0.2)
    assert abs(opt.rate(1000) - 0.2) < 1e-12


def _sampling_worker(rank, world, port, q):
    """Ranks with DIFFERENT torch seeds must still run the same supernet path: the step harness broadcasts rank 0's
    picks (the reference only relies on identically seeded ranks, search_vqa.py:62-66)."""
    import numpy as np
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from mmnas_b200.data.synthetic import Cfg
        from mmnas_b200.model.nets import Net_Search
        cfg = Cfg(mode='search', HSIZE=64, FRCNFEAT_SIZE=32, BBOXFEAT_EMB_SIZE=32, WORD_EMBED_SIZE=16, ATTFLAT_MLP_SIZE=48,
                  ATTFLAT_OUT_SIZE=128)
        torch.manual_seed(888)
        net = Net_Search(cfg, {'token_size': 30, 'ans_size': 11, 'pretrained_emb': np.zeros((30, 16), np.float32)})
        torch.manual_seed(888 + 1000 * rank)               # ranks drift apart on purpose
        draws = []
        for _ in range(3):
            net.reset_binary_gates(batched=True)
            draws.append([m.active_index[0] for m in net.redundant_modules])
            for m in net.redundant_modules:
                assert float(m.alpha_gate.data[m.active_index[0]]) == 1.0 and float(m.alpha_gate.data.sum()) == 1.0
        q.put((rank, draws))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_search_sampling_is_rank_invariant_under_data_parallelism():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_sampling_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert got[0] == got[1]
    from tests.util import load_golden
    assert got[0] == load_golden('sampling_seed888.npz')['draws_seed888'].tolist()[:3]     # rank 0 draws the seed-888 path
