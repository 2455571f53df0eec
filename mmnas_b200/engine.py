"""Step harness around the operator hot path: the train step of train_vqa.py:294-311 and the search step of
search_vqa.py:278-337 on synthetic batches, data-parallel over one process per GPU.

Pieces
  FlatGrads      all gradients live in one flat fp32 buffer (views installed as p.grad); zeroing it is one memset
                 and gives every parameter a (zero) gradient each step — what the reference obtains with its
                 `loss += 0 * sum(p.sum())` trick (train_vqa.py:298, search_vqa.py:286-288).
  BucketReducer  data-parallel gradient averaging: the flat buffer is cut into buckets in reverse registration
                 order (≈ backward order); a bucket's all-reduce is launched from the post-accumulate hook of its
                 last parameter, so NCCL traffic over NVLink overlaps the rest of the backward.  Replaces DDP's
                 reducer (train_vqa.py:236); same semantics: mean over ranks of per-rank sum-reduced losses.
  WarmupAdam     Adam + the 1/4,2/4,3/4,1 warm-up of mmnas/utils/optimizer.py:25-44 (+ decay), clip_grad_norm_.
  TrainStep / SearchStep   the step bodies; TrainStep can replay the whole step as one CUDA graph.
  Prefetcher     pinned host batches -> device on a copy stream, double buffered.
"""
import os

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import runtime
from .model.mixed import MixedOp


# ------------------------------------------------------------------------------------------------ gradients
class FlatGrads:
    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        dev = self.params[0].device
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4          # keep every view 16-byte aligned
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]
        self.install()

    def view(self, i):
        return self.views[i]

    def install(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        self.flat.zero_()
        for p, v in zip(self.params, self.views):   # MixedOp.binarize() sets candidate grads to None
            if p.grad is not v:
                p.grad = v


class BucketReducer:
    """Overlapped gradient mean over the process group, on top of FlatGrads."""

    def __init__(self, flat_grads, group=None, bucket_mb=60.0):
        self.fg = flat_grads
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.enabled = self.world > 1
        self.buckets = []          # (start, end, [param indices]) over the flat buffer, reverse registration order
        bucket_mb = float(os.environ.get('MMNAS_BUCKET_MB', bucket_mb))      # tuning / A-B experiments only
        cap = int(bucket_mb * (1 << 20) / 4)
        idxs, end = [], None
        n = len(self.fg.params)
        for i in reversed(range(n)):
            p_end = self.fg.offsets[i] + (self.fg.params[i].numel() + 3) // 4 * 4
            if end is None:
                end = p_end
            idxs.append(i)
            if end - self.fg.offsets[i] >= cap or i == 0:
                self.buckets.append((self.fg.offsets[i], end, idxs))
                idxs, end = [], None
        self.bucket_of = {}
        for b, (_, _, ids) in enumerate(self.buckets):
            for i in ids:
                self.bucket_of[i] = b
        self._pending = [0] * len(self.buckets)
        self._fired = [False] * len(self.fg.params)
        self._launched = [False] * len(self.buckets)
        self._works = []
        self._avg = None
        self._index = {id(p): i for i, p in enumerate(self.fg.params)}
        self._hooks = [self._make_hook(i) for i in range(len(self.fg.params))]
        self._armed = False
        if self.enabled:
            for i, p in enumerate(self.fg.params):
                p.register_post_accumulate_grad_hook(self._hooks[i])
        self.reset()

    def _make_hook(self, i):
        def hook(param):
            if not self._armed or self._fired[i]:
                return                      # (autograd also fires the hook for a gradient returned as None,
            self._fired[i] = True           #  after notify() already counted the directly written one)
            b = self.bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def notify(self, param):
        """Same as the autograd hook, for gradients a block backward accumulated straight into p.grad."""
        i = self._index.get(id(param))
        if i is not None and self.enabled:
            self._hooks[i](param)

    def reset(self):
        """Arm for one backward pass."""
        for b, (_, _, ids) in enumerate(self.buckets):
            self._pending[b] = len(ids)
            self._launched[b] = False
        self._works = []
        self._fired = [False] * len(self.fg.params)
        self._armed = True

    def _launch(self, b):
        if self._launched[b] or not self.enabled:
            return
        self._launched[b] = True
        s, e, _ = self.buckets[b]
        chunk = self.fg.flat[s:e]
        if self._avg is None:
            self._avg = dist.get_backend(self.group) == 'nccl'
        if self._avg:
            self._works.append(dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:                                       # gloo has no AVG
            chunk.mul_(1.0 / self.world)
            self._works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Launch buckets whose parameters never fired (unused candidates), then wait for everything."""
        self._armed = False
        if not self.enabled:
            return
        for b in range(len(self.buckets)):
            self._launch(b)
        for w in self._works:
            w.wait()
        self._works = []


class WeightShadows:
    """bf16 copies of every GEMM weight of the net (stacked per fused GEMM), refreshed by ONE batched cast kernel
    per step instead of ~94 small casts.  Buffers are registered with their owner modules as 'managed'; the
    modules trust them only while `runtime.shadows_fresh` is set (inside an engine step)."""
    CHUNK = 4096

    def __init__(self, net):
        from . import kernels
        self.kernels = kernels
        rows = []
        self.buffers = []
        for mod in net.modules():
            if not hasattr(mod, 'shadow_specs'):
                continue
            for owner, name, params in mod.shadow_specs():
                if not params[0].is_cuda:
                    continue
                cols = params[0].shape[1]
                buf = torch.empty((sum(p.shape[0] for p in params), cols), dtype=torch.bfloat16, device=params[0].device)
                owner._w16[name] = ('managed', buf)
                self.buffers.append(buf)
                off = 0
                for p in params:
                    n = p.numel()
                    assert n % 4 == 0 and p.is_contiguous()
                    for c in range(0, n, self.CHUNK):
                        rows.append((p.data_ptr() + 4 * c, buf.data_ptr() + 2 * (off + c), min(self.CHUNK, n - c)))
                    off += n
        self.n_chunks = len(rows)
        self.table = torch.tensor(rows, dtype=torch.int64, device=self.buffers[0].device) if rows else None
        self._watch = [(p, p.data_ptr()) for p in self._watched(net)]

    @staticmethod
    def _watched(net):
        ps = [p for mod in net.modules() if hasattr(mod, 'shadow_specs') for _, _, params in mod.shadow_specs()
              for p in params if p.is_cuda]
        return ps[:1] + ps[-1:]

    def refresh(self):
        # the device table holds raw parameter addresses: a storage change (net.to(), p.data = ...) would make the
        # kernel read stale memory without any error, so the first and last entries are re-checked on every call
        for p, ptr in self._watch:
            if p.data_ptr() != ptr:
                raise RuntimeError('parameter storage moved after the engine step was built (net.to() / p.data = ...): '
                                   're-create the TrainStep / SearchStep')
        if self.n_chunks:
            self.kernels.cast_multi(self.table, self.n_chunks)


# ------------------------------------------------------------------------------------------------ optimizer
class WarmupAdam:
    """Adam(lr schedule of WarmupOptimizer, optimizer.py:14-47) preceded by clip_grad_norm_ (train_vqa.py:309-311)."""

    def __init__(self, params, lr_base, epoch_steps, betas=(0.9, 0.98), eps=1e-9, clip=1.0, warmup=True,
                 capturable=False, flat_grads=None):
        self.params = [p for p in params if p.requires_grad]
        self.lr_base, self.epoch_steps, self.warmup, self.clip = lr_base, max(1, epoch_steps), warmup, clip
        self._step = 0
        self._rate = 0.0
        self.capturable = capturable
        self.fused = None
        if flat_grads is not None and self.params[0].is_cuda:
            # clip + Adam as two CUDA passes over the flat gradient buffer (mmnas_sumsq_f32 + mmnas_clip_adam)
            self.fused = FusedClipAdam(flat_grads, self.params, betas, eps, clip)
            self.optimizer = None
            return
        lr = torch.tensor(0.0, device=self.params[0].device) if capturable else 0.0
        self.optimizer = torch.optim.Adam(self.params, lr=lr, betas=betas, eps=eps, weight_decay=0,
                                          fused=self.params[0].is_cuda, capturable=capturable)

    def rate(self, step=None):
        step = self._step if step is None else step
        if not self.warmup:
            return self.lr_base
        for k in (1, 2, 3):
            if step <= int(self.epoch_steps * k):
                return self.lr_base * k / 4.
        return self.lr_base

    def decay(self, r):
        self.lr_base *= r

    def set_start_step(self, step):
        """optimizer.py:48-49 (resume): the schedule continues from `step`."""
        self._step = step

    def state_dict(self):
        return self.fused.state_dict() if self.fused is not None else self.optimizer.state_dict()

    def load_state_dict(self, sd):
        (self.fused if self.fused is not None else self.optimizer).load_state_dict(sd)

    def set_lr(self):
        """Host side of a step: advance the schedule and publish the learning rate."""
        self._step += 1
        self._rate = self.rate()
        if self.fused is not None:
            self.fused.lr.fill_(self._rate)
            return
        for g in self.optimizer.param_groups:
            if torch.is_tensor(g['lr']):
                g['lr'].fill_(self._rate)
            else:
                g['lr'] = self._rate

    def clip_and_step(self):
        if self.fused is not None:
            self.fused.step()
            return
        if self.clip and self.clip > 0:
            torch.nn.utils.clip_grad_norm_(self.params, self.clip, foreach=True)
        self.optimizer.step()


class FusedClipAdam:
    """clip_grad_norm_(max_norm) + Adam over the parameters `params`, whose gradients are a CONTIGUOUS prefix range of
    a FlatGrads buffer: one reduction pass over the flat gradients, one tabled update pass over (param, grad, m, v).
    Learning rate and step count live on the device, so the step replays from a CUDA graph."""
    CHUNK = 4096

    def __init__(self, flat_grads, params, betas, eps, clip):
        from . import kernels
        self.kernels = kernels
        fg = flat_grads
        wanted = {id(q) for q in params}
        idx = [i for i, p in enumerate(fg.params) if id(p) in wanted]
        assert idx == list(range(idx[0], idx[0] + len(idx))), 'parameters must be contiguous in the flat gradient buffer'
        first, last = idx[0], idx[-1]
        self.lo = fg.offsets[first]
        self.hi = fg.offsets[last] + (fg.params[last].numel() + 3) // 4 * 4
        dev = fg.flat.device
        self.grad_range = fg.flat[self.lo:self.hi]
        self.exp_avg = torch.zeros(self.hi - self.lo, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.hi - self.lo, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.scratch = torch.zeros(kernels.SUMSQ_SCRATCH, dtype=torch.float32, device=dev)
        self.lr = torch.zeros(1, dtype=torch.float32, device=dev)
        self.state = torch.zeros(2, dtype=torch.int64, device=dev)        # [unused, step count]
        self.betas, self.eps, self.clip = betas, eps, (clip if clip else 0.0)
        rows = []
        for i in idx:
            p = fg.params[i]
            off = fg.offsets[i] - self.lo
            g_ptr = fg.flat.data_ptr() + 4 * fg.offsets[i]
            for c in range(0, p.numel(), self.CHUNK):
                rows.append((p.data_ptr() + 4 * c, g_ptr + 4 * c, self.exp_avg.data_ptr() + 4 * (off + c),
                             self.exp_avg_sq.data_ptr() + 4 * (off + c), min(self.CHUNK, p.numel() - c)))
            assert p.is_contiguous()
        self.n_chunks = len(rows)
        self.table = torch.tensor(rows, dtype=torch.int64, device=dev)
        self.params = [fg.params[i] for i in idx]
        self.offsets = [fg.offsets[i] - self.lo for i in idx]
        self._watch = [(fg.params[i], fg.params[i].data_ptr(), fg.views[i].data_ptr(), i) for i in (first, last)]
        self._fg = fg

    def _check_storage(self):
        for p, ptr, gptr, i in self._watch:
            if p.data_ptr() != ptr or self._fg.views[i].data_ptr() != gptr:
                raise RuntimeError('parameter / gradient storage moved after the optimizer was built (net.to(), '
                                   'p.data = ..., a new FlatGrads): re-create the TrainStep / SearchStep')

    # torch.optim.Adam's state layout, so that the reference checkpoint entry 'net_optim' (train_vqa.py:315-321) can be
    # produced and resumed: {'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]}
    def state_dict(self):
        step = float(self.state[1].item())
        st = {}
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            n = p.numel()
            st[i] = {'step': torch.tensor(step), 'exp_avg': self.exp_avg[off:off + n].view_as(p).clone(),
                     'exp_avg_sq': self.exp_avg_sq[off:off + n].view_as(p).clone()}
        return {'state': st, 'param_groups': [{'lr': float(self.lr.item()), 'betas': tuple(self.betas), 'eps': self.eps,
                                               'weight_decay': 0, 'amsgrad': False,
                                               'params': list(range(len(self.params)))}]}

    def load_state_dict(self, sd):
        step = 0
        with torch.no_grad():
            for i, (p, off) in enumerate(zip(self.params, self.offsets)):
                ent = sd['state'].get(i)
                if ent is None:
                    continue
                n = p.numel()
                self.exp_avg[off:off + n].copy_(ent['exp_avg'].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(ent['exp_avg_sq'].reshape(-1))
                step = max(step, int(float(ent['step'])))
            self.state[1] = step

    def step(self):
        self._check_storage()
        k = self.kernels
        k.rng_advance(self.state)
        if self.clip > 0:
            k.sumsq(self.grad_range, self.sumsq, self.scratch)
        k.clip_adam(self.table, self.n_chunks, self.sumsq, self.lr, self.state, self.betas[0], self.betas[1], self.eps,
                    float(self.clip))


# ------------------------------------------------------------------------------------------------ steps
def vqa_loss(pred, target):
    """train_vqa.py:297: BCEWithLogitsLoss(reduction='sum')."""
    return F.binary_cross_entropy_with_logits(pred, target, reduction='sum')


def vgd_loss(pred, target, loss_lambda=0.5):
    """train_vgd.py:320-334 with the shipped settings (SCORES_LOSS 'kld', LOSS_AVG True, LOSS_LAMBDA 0.5, REDUCTION
    'sum'): KLDivLoss over the masked log-softmax region scores / #scored samples + 0.5 * SmoothL1 over the masked box
    regressions / #matching regions.  target = (scores, scores_mask, transformed_bbox, bbox_mask)."""
    pred_scores, pred_reg = pred
    scores, scores_mask, tbox, bbox_mask = target
    loss_scores = F.kl_div(pred_scores * scores_mask, scores * scores_mask, reduction='sum')
    loss_reg = F.smooth_l1_loss(pred_reg * bbox_mask, tbox * bbox_mask, reduction='sum')
    return loss_scores / scores_mask.sum() + loss_lambda * (loss_reg / bbox_mask.sum())


def itm_loss(pred, target=None):
    """mmnas/utils/itm_loss.py:13-24 (BCE_Loss, reduction 'sum'): loss_pos + loss_negc + loss_pos + loss_negi — the
    positive term is counted twice, as in the reference.  `pred` holds the sigmoid scores of the three forwards of
    train_itm.py:387-389 stacked along the batch: [positive | negative captions | negative images]."""
    pos, negc, negi = pred.chunk(3, dim=0)
    one, zero = torch.ones_like(pos), torch.zeros_like(pos)
    bce = F.binary_cross_entropy
    return 2.0 * bce(pos, one, reduction='sum') + bce(negc, zero, reduction='sum') + bce(negi, zero, reduction='sum')


LOSSES = {'vqa': vqa_loss, 'vgd': vgd_loss, 'itm': itm_loss}


def tree_map(fn, obj):
    """Apply fn to every tensor of a (nested) tuple / list; None passes through."""
    if obj is None:
        return None
    if torch.is_tensor(obj):
        return fn(obj)
    return tuple(tree_map(fn, o) for o in obj)


def tree_leaves(obj):
    if obj is None:
        return []
    if torch.is_tensor(obj):
        return [obj]
    return [t for o in obj for t in tree_leaves(o)]


class TrainStep:
    """net_optim.zero_grad(); pred = net(x); loss = BCE_sum; backward (+ overlapped grad mean); clip; Adam."""

    def __init__(self, net, lr_base=0.00012, epoch_steps=1000, bucket_mb=60.0, use_graph=False, loss_fn=vqa_loss,
                 betas=(0.9, 0.98), eps=1e-9, clip=1.0):
        self.net = net
        self.loss_fn = loss_fn
        self.grads = FlatGrads(net.parameters())
        self.reducer = BucketReducer(self.grads, bucket_mb=bucket_mb)
        # NCCL all-reduces launched from the backward hooks are captured with the rest of the step
        self.use_graph = use_graph
        self.optim = WarmupAdam(self.grads.params, lr_base, epoch_steps, betas, eps, clip, capturable=self.use_graph,
                                flat_grads=self.grads)
        self.shadows = WeightShadows(net)
        self.graph = None
        self._static_in = self._static_tgt = self._static_loss = None

    def _body(self, inputs, target):
        self.grads.zero()
        self.reducer.reset()
        runtime.advance(self.grads.flat.device)
        if runtime.get_precision() == 'bf16':
            self.shadows.refresh()
            runtime.shadows_fresh = True
        runtime.direct_grads, runtime.grad_listener = True, (self.reducer.notify if self.reducer.enabled else None)
        try:
            pred = self.net(inputs)
            loss = self.loss_fn(pred, target)
            loss.backward()
        finally:
            runtime.shadows_fresh = False
            runtime.direct_grads, runtime.grad_listener = False, None
        self.reducer.finish()
        self.optim.clip_and_step()
        return loss.detach()

    def __call__(self, inputs, target):
        self.optim.set_lr()
        if not self.use_graph:
            return self._body(inputs, target)
        if self.graph is None:
            self._capture(inputs, target)
        for dst, src in zip(tree_leaves(self._static_in) + tree_leaves(self._static_tgt),
                            tree_leaves(inputs) + tree_leaves(target)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self._static_loss

    def _mutable_state(self):
        """Every tensor a step mutates: parameters, Adam moments + step counter, the dropout step counter.  (Gradients
        and bf16 shadows are rewritten from scratch by each step.)"""
        st = list(self.grads.params)
        f = self.optim.fused
        st += [f.exp_avg, f.exp_avg_sq, f.state]
        st.append(runtime.rng_state(self.grads.flat.device))
        return st

    def _capture(self, inputs, target):
        """Warm up (allocator pools, lazy initialisation) and capture the step.  The warm-up steps run on the first
        batch; everything they mutate is snapshotted before and restored after, so the FIRST replay is the first
        update — graph and eager runs evolve identically (parameters, Adam moments, bias-correction count, dropout
        step counter)."""
        self._static_in = tree_map(torch.clone, inputs)
        self._static_tgt = tree_map(torch.clone, target)
        if self.optim.fused is None:
            raise RuntimeError('graph capture needs CUDA parameters (fused clip + Adam)')
        state = self._mutable_state()
        saved = [t.detach().clone() for t in state]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):               # warm-up outside capture
            for _ in range(3):
                self._body(self._static_in, self._static_tgt)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._static_loss = self._body(self._static_in, self._static_tgt)
        with torch.no_grad():                       # capture executes nothing, the warm-up did: undo it
            for t, v in zip(state, saved):
                t.copy_(v)
        torch.cuda.synchronize()

    @property
    def static_inputs(self):
        return self._static_in, self._static_tgt


class _Segments:
    """State of SearchStep's segmented replay: static inputs, the tensors handed from one segment to the next, the four
    captured graphs."""
    key = ex_key = inputs = target = x_in = y_in = x_mask = y_mask = x_out = y_out = loss = graphs = graph_launches = None


class SearchStep:
    """One iteration of search_vqa.py:278-337: weight step on the sampled path, and (when `arch=True`) the
    architecture step in MODE 'full' on a held-out batch."""

    def __init__(self, net, lr_base=0.0004, epoch_steps=1000, alpha_lr=0.1, alpha_betas=(0., 0.999), mode='full',
                 bucket_mb=60.0, loss_fn=vqa_loss, use_executor=True, segments=True):
        self.net = net
        self.mode = mode
        self.loss_fn = loss_fn
        self.net_params = [p for p in net.net_parameters()]
        alphas = list(net.alpha_prob_parameters()) + list(net.alpha_gate_parameters())
        self.grads = FlatGrads(self.net_params + alphas)  # weights first (a contiguous range for clip + Adam), then alphas
        self.reducer = BucketReducer(self.grads, bucket_mb=bucket_mb)
        self.optim = WarmupAdam(self.net_params, lr_base, epoch_steps, flat_grads=self.grads)
        self.alpha_optim = torch.optim.Adam(list(net.alpha_prob_parameters()), alpha_lr, betas=alpha_betas,
                                            weight_decay=0)
        self.shadows = WeightShadows(net)       # every candidate's GEMM weights, one batched cast per step
        # static-plan executor of the supernet backbone (executor.py): one autograd node, one foreign call per block
        self.executor = None
        if use_executor and self.net_params[0].is_cuda and mode == 'full':
            from .executor import SearchExecutor
            self.executor = SearchExecutor(net)
        object.__setattr__(net, '_executor', self.executor)
        # Segmented replay (see _Segments): everything around the sampled backbone is static, so it replays from four
        # CUDA graphs while the backbone stays a sequence of block-level calls chosen per step.
        self.presample = self.net_params[0].is_cuda and os.environ.get('MMNAS_PRESAMPLE', '1') != '0'
        self.use_segments = (segments and self.executor is not None and getattr(net, 'rel_mode', 'geometry') == 'geometry'
                             and os.environ.get('MMNAS_SEARCH_SEGMENTS', '1') != '0')
        self._seg = None

    def _forward_backward(self, inputs, target):
        if self.use_segments:
            return self._segmented(inputs, target)
        self.grads.zero()                # net.zero_grad() + the reference's 0*sum(params) dummy terms
        self.reducer.reset()
        runtime.advance(self.grads.flat.device)
        if runtime.get_precision() == 'bf16':
            self.shadows.refresh()
            runtime.shadows_fresh = True
        runtime.direct_grads, runtime.grad_listener = True, (self.reducer.notify if self.reducer.enabled else None)
        try:
            loss = self.loss_fn(self.net(inputs), target)
            loss.backward()
        finally:
            runtime.shadows_fresh = False
            runtime.direct_grads, runtime.grad_listener = False, None
        self.reducer.finish()
        return loss.detach()

    # ---- segmented replay ------------------------------------------------------------------------------------------
    # A: zero gradients, advance the dropout counter, refresh the bf16 weight shadows, stem forward (embedding, LSTM,
    #    region projection, masks, box geometry), stage the executor's inputs
    #    -> executor.run_forward(): eager block-level calls along the sampled path (or all candidates, MODE 'full')
    # B: heads + loss forward and backward, output gradients staged for the executor
    #    -> executor.run_backward()
    # C: stem backward (region projection, LSTM, embedding)
    #    -> gradient mean over ranks (eager NCCL, world > 1)
    # D: clip + Adam (weight step only)
    # A and C share one autograd graph, exactly as torch.cuda.make_graphed_callables splits forward and backward.
    def _seg_a(self, seg):
        self.grads.zero()
        runtime.advance(self.grads.flat.device)
        if runtime.get_precision() == 'bf16':
            self.shadows.refresh()
        x_in, y_in, x_mask, y_mask, _, g4 = self.net.stem(seg.inputs)
        ex = self.executor
        ex.prepare(x_in, y_in)
        ex.stage(x_in, y_in, x_mask, y_mask, g4)
        seg.x_in, seg.y_in, seg.x_mask, seg.y_mask = x_in, y_in, x_mask, y_mask

    def _seg_b(self, seg):
        ex = self.executor
        if seg.x_out is None or seg.x_out.data_ptr() != ex.xs[-1].data_ptr():
            seg.x_out = ex.xs[-1].detach().requires_grad_(True)      # leaves aliasing the executor's output buffers
            seg.y_out = ex.ys[-1].detach().requires_grad_(True)
        seg.x_out.grad = seg.y_out.grad = None
        loss = self.loss_fn(self.net.head(seg.x_out, seg.y_out, seg.x_mask, seg.y_mask), seg.target)
        loss.backward()
        for dst, leaf in ((ex.dxs[-1], seg.x_out), (ex.dys[-1], seg.y_out)):
            if leaf.grad is None:
                dst.zero_()
            else:
                dst.copy_(leaf.grad)
        seg.x_out.grad = seg.y_out.grad = None
        seg.loss = loss.detach()

    def _seg_c(self, seg):
        ex = self.executor
        torch.autograd.backward([seg.x_in, seg.y_in], [ex.dxs[0], ex.dys[0]])
        seg.x_in = seg.y_in = None

    def _seg_d(self, seg):
        self.optim.clip_and_step()

    def _reduce(self):
        if self.reducer.enabled:
            self.reducer.reset()
            self.reducer.finish()

    def _seg_key(self, inputs, target):
        return (tuple((tuple(t.shape), t.dtype) for t in tree_leaves(inputs) + tree_leaves(target)),
                runtime.get_precision(), self.net.training)

    def _capture_segments(self, inputs, target):
        """Warm up eagerly (plans, allocator pools, cuDNN handles), then capture A-D into one memory pool.  Whatever
        the warm-up mutates is restored, so the first replayed step is the first update."""
        seg = _Segments()
        seg.key = self._seg_key(inputs, target)
        seg.inputs = tree_map(torch.clone, inputs)
        seg.target = tree_map(torch.clone, target)
        ex = self.executor
        f = self.optim.fused
        state = list(self.net_params) + [f.exp_avg, f.exp_avg_sq, f.state, runtime.rng_state(self.grads.flat.device)]
        saved = [t.detach().clone() for t in state]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._seg_a(seg)
                ex.run_forward()
                self._seg_b(seg)
                ex.run_backward()
                self._seg_c(seg)
                self._seg_d(seg)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        pool = torch.cuda.graph_pool_handle()
        seg.graphs, seg.graph_launches = [], []
        from . import _lib
        for body in (self._seg_a, self._seg_b, self._seg_c, self._seg_d):
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launches()
            with torch.cuda.graph(g, pool=pool):
                body(seg)
            seg.graphs.append(g)
            seg.graph_launches.append(_lib.launches() - l0)      # library kernels inside this graph (replayed, not re-counted)
        seg.ex_key = ex.key
        with torch.no_grad():
            for t, v in zip(state, saved):
                t.copy_(v)
        torch.cuda.synchronize()
        self._seg = seg

    def _segmented(self, inputs, target):
        seg = self._seg
        self.reducer._armed = False          # no per-parameter hooks: the mean over ranks runs once, after segment C
        runtime.direct_grads, runtime.grad_listener = True, None
        runtime.shadows_fresh = runtime.get_precision() == 'bf16'
        try:
            if seg is None or seg.key != self._seg_key(inputs, target) or seg.ex_key != self.executor.key:
                self._seg = None
                self._capture_segments(inputs, target)
                seg = self._seg
            for dst, src in zip(tree_leaves(seg.inputs) + tree_leaves(seg.target),
                                tree_leaves(inputs) + tree_leaves(target)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            ga, gb, gc, gd = seg.graphs
            ex = self.executor
            ga.replay()
            ex.run_forward()
            gb.replay()
            ex.run_backward()
            gc.replay()
            self._reduce()
        finally:
            runtime.shadows_fresh = False
            runtime.direct_grads, runtime.grad_listener = False, None
        return seg.loss.clone()

    def replayed_launches(self, arch=False):
        """Library kernels a step replays from its captured segments (not seen by the library's launch counter)."""
        if self._seg is None:
            return 0
        n = self._seg.graph_launches
        return sum(n[:3]) + (0 if arch else n[3])

    def _off(self):
        # Net_Search.unused_modules_off/back (hygr_vqa.py:175-196) swap the candidates that do not run for None; the
        # executor addresses candidates through its static plans and never touches the others, so the 130 module-list
        # writes per step are skipped when it is active
        if self.executor is None:
            self.net.unused_modules_off()

    def _back(self):
        if self.executor is None:
            self.net.unused_modules_back()

    def weight_step(self, inputs, target):
        MixedOp.MODE = None
        self.net.reset_binary_gates(batched=True)
        self._off()
        try:
            self.optim.set_lr()
            loss = self._forward_backward(inputs, target)
            if self.use_segments:
                self._seg.graphs[3].replay()
            else:
                self.optim.clip_and_step()
        finally:
            self._back()
        self._presample()
        return loss

    def arch_step(self, inputs, target):
        MixedOp.MODE = self.mode
        self.net.reset_binary_gates(batched=(self.mode != 'two'))
        self._off()
        try:
            loss = self._forward_backward(inputs, target)
            self.net.set_arch_param_grad(batched=True)
            self.alpha_optim.step()
            if MixedOp.MODE == 'two':
                self.net.rescale_updated_arch_param()
            self.net.alphas_updated()
        finally:
            self._back()
            MixedOp.MODE = None
        self._presample()
        return loss

    def _presample(self):
        # next step's draw, issued now on the sampling stream (nets.py: presample); MODE 'two' samples per module
        if self.presample and self.mode != 'two':
            self.net.presample()

    def __call__(self, train_batch, eval_batch=None):
        loss = self.weight_step(*train_batch)
        if eval_batch is not None:
            loss = self.arch_step(*eval_batch)
        return loss


# ------------------------------------------------------------------------------------------------ input pipeline
class Prefetcher:
    """Host (pinned) -> device copies on a side stream, one batch ahead of the compute stream."""

    def __init__(self, host_batches, device):
        self.device = torch.device(device)
        self.host = [self._pin(b) for b in host_batches]
        self.stream = torch.cuda.Stream(device=self.device)
        self.bytes_per_batch = sum(t.numel() * t.element_size() for t in self._flat(self.host[0]))
        self._i = 0
        self._next = None
        self._issue()

    @staticmethod
    def _flat(batch):
        return tree_leaves(batch)

    def _pin(self, batch):
        return tree_map(lambda t: t.pin_memory(), batch)

    def _issue(self):
        inputs, target = self.host[self._i % len(self.host)]
        self._i += 1
        with torch.cuda.stream(self.stream):
            dev_in = tree_map(lambda t: t.to(self.device, non_blocking=True), inputs)
            dev_tgt = tree_map(lambda t: t.to(self.device, non_blocking=True), target)
        self._next = (dev_in, dev_tgt)

    def next(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        batch = self._next
        for t in self._flat(batch):
            t.record_stream(torch.cuda.current_stream(self.device))
        self._issue()
        return batch


class ScalarLog:
    """Per-step scalars (the loss) to the host without stalling the launch queue: push() enqueues a non-blocking
    device->host copy into pinned memory behind the step that produced the value, pop() returns the oldest pushed value
    (waiting for its copy only).  Reading step t's loss while step t+1 is already queued keeps the GPU busy; a
    synchronous float(loss) after every step leaves it idle for the host's launch latency each time."""

    def __init__(self, device, depth=4):
        self.device = torch.device(device)
        self.slots = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.head = self.tail = 0

    def __len__(self):
        return self.head - self.tail

    def push(self, value):
        if len(self) == len(self.slots):
            raise RuntimeError('ScalarLog full: pop() before pushing more')
        i = self.head % len(self.slots)
        self.slots[i].copy_(value.detach().reshape(1), non_blocking=True)
        self.events[i].record(torch.cuda.current_stream(self.device))
        self.head += 1

    def pop(self):
        if not len(self):
            raise RuntimeError('ScalarLog empty')
        i = self.tail % len(self.slots)
        self.events[i].synchronize()
        self.tail += 1
        return float(self.slots[i][0])
