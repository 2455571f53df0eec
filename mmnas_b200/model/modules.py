"""Candidate blocks of the MMnas search space, B200-native.

Drop-in for the reference's mmnas/model/modules.py on the operator hot path: same class names, constructor
signatures (`Op(__C, norm=False, residual=False, ...)`), `forward(x, y, x_mask, y_mask, rel_embed)` call
convention and state-dict keys (`mhatt.linear_{v,k,q,merge}.weight`, `mhatt.linear_r.{weight,bias}`,
`mlp.fc.linear.*`, `mlp.linear.*`, `ln.{a_2,b_2}`), so reference checkpoints load unchanged.  The arithmetic
of SelfAtt / RelSelfAtt / GuidedAtt / FeedForward / LayerNorm runs in libmmnas_b200 (hand-written sm_100a
CUDA); the nn.Linear sub-modules here are parameter containers only.  Without the library, or on CPU
tensors, forward raises — there is no fallback.

Reference lines: LayerNorm modules.py:44-56, FC :13-31, MLP :34-41, MHAtt :158-199, RelMHAtt :202-245,
SelfAtt :248-271, RelSelfAtt :274-298, GuidedAtt :301-325, FeedForward :328-362, AttFlat :59-85.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import kernels as K
from .. import runtime
from .._lib import require_cuda
from ..functional import (AttBlockFn, AttBlockPyFn, FFNBlockFn, FFNBlockPyFn, LayerNormFn, BlockCfg, LinearFn,
                          AddLayerNormFn)


class RelGeometry:
    """Lazy handle for the RSA geometry embedding: the raw 4-d box geometry plus the `linear_y_rel` layer
    (full_vqa.py:82,103).  Handing this to RelSelfAtt instead of the dense relu(linear_y_rel(g)) tensor lets
    the kernel rebuild the 64-d embedding on the fly, so [B,N,N,REL_SIZE] never reaches HBM; gradients of
    the layer are produced by the RSA backward kernels."""

    def __init__(self, g4, linear):
        self.g4 = g4
        self.weight = linear.weight
        self.bias = linear.bias
        # every RSA block of the net contributes to these two gradients: keep them on autograd's accumulation path so
        # the data-parallel reducer sees ONE completion per step (see functional._grad_target)
        self.weight._mmnas_shared = True
        self.bias._mmnas_shared = True

    def dense(self):
        return F.relu(F.linear(self.g4, self.weight, self.bias))


def _key_mask(mask, B, N):
    if mask is None:
        return None
    m = mask.reshape(B, -1)
    if m.shape[1] != N:
        raise ValueError('attention mask has %d keys, expected %d' % (m.shape[1], N))
    m = m.contiguous()
    return m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)


def _shadow(t):
    s = getattr(t, '_mmnas_bf16', None)
    if s is not None and s.shape == t.shape and s.device == t.device:
        return s
    return None


class _OpBase(nn.Module):
    """Shared plumbing: dropout sites, bf16 weight shadows, precision selection."""

    def _init_runtime(self, n_sites):
        self._sites = [runtime.new_site() for _ in range(n_sites)]
        self._calls = 0
        self._epoch = -1
        self._w16 = {}
        self.precision = None      # None -> follow mmnas_b200.set_precision()

    def _mode(self):
        return self.precision or runtime.get_precision()

    def _drops(self, x, p):
        if not (self.training and p > 0):
            return tuple(K.NO_DROP for _ in self._sites)
        if self._epoch != runtime.epoch:        # calls are numbered within a step (runtime.call_index)
            self._epoch, self._calls = runtime.epoch, 0
        self._calls += 1
        st = runtime.rng_state(x.device)
        return tuple(K.Drop(st, (s << 32) | (self._calls & 0xFFFFFFFF), p) for s in self._sites)

    def _fused_bf16(self, name, params):
        """bf16 copy of `params` stacked along dim 0.  Re-cast on every training forward (the optimizer step may
        sit inside a captured CUDA graph, where version counters do not move); version-checked in eval."""
        ent = self._w16.get(name)
        if ent is not None and ent[0] == 'managed' and runtime.shadows_fresh:
            return ent[1]            # refreshed by engine.WeightShadows in one batched cast this step
        if ent is not None and ent[0] == 'managed':
            ent = (None, ent[1])     # outside an engine step: cast here, into the same buffer
        vers = tuple((p.data_ptr(), p._version) for p in params)
        live = self.training
        if ent is not None and not live and ent[0] == vers and ent[1].device == params[0].device:
            return ent[1]
        rows = sum(p.shape[0] for p in params)
        cols = params[0].shape[1]
        buf = ent[1] if (ent is not None and ent[1].shape == (rows, cols) and ent[1].device == params[0].device) \
            else torch.empty((rows, cols), dtype=torch.bfloat16, device=params[0].device)
        r = 0
        for p in params:
            K.cast_bf16(p.detach(), buf[r:r + p.shape[0]])
            r += p.shape[0]
        if self._w16.get(name, (None,))[0] != 'managed':
            self._w16[name] = (vers, buf)
        return buf


class LayerNorm(nn.Module):
    def __init__(self, size, eps=1e-6, dim=-1):
        super().__init__()
        self.eps = eps
        self.dim = dim
        self.a_2 = nn.Parameter(torch.ones(size))
        self.b_2 = nn.Parameter(torch.zeros(size))

    def forward(self, x):
        if self.dim not in (-1, x.dim() - 1):
            raise NotImplementedError('LayerNorm is implemented over the last dimension only')
        return LayerNormFn.apply(x, self.a_2, self.b_2, self.eps)


def linear(x, lin, relu=False, drop=K.NO_DROP):
    """nn.Linear `lin` (+ReLU, + dropout) on the library GEMMs for CUDA tensors (LinearFn); plain torch for CPU tensors —
    these layers belong to the callers (stem / heads), not to the operator hot path, so a CPU caller
    (constructing or inspecting a net) is not an error."""
    if x.is_cuda:
        return LinearFn.apply(x, lin.weight, lin.bias, relu, drop, runtime.get_precision())
    y = lin(x)
    return F.relu(y) if relu else y


class FC(nn.Module):
    """Linear(+ReLU+dropout), modules.py:13-31.  Inside FeedForward the arithmetic is fused into the block kernels (this
    module is then a parameter container); the stand-alone forward serves the task heads (AttFlat) and runs on the
    library GEMM with the bias / ReLU / dropout epilogue."""

    def __init__(self, in_size, out_size, dropout_r=0., use_relu=True):
        super().__init__()
        self.dropout_r = dropout_r
        self.use_relu = use_relu
        self.linear = nn.Linear(in_size, out_size)
        self._site = runtime.new_site()
        self._ncalls = [0]
        self._epoch = [-1]

    def forward(self, x):
        drop = K.NO_DROP
        if self.training and self.dropout_r > 0 and self.use_relu and x.is_cuda:
            if self._epoch[0] != runtime.epoch:
                self._epoch[0], self._ncalls[0] = runtime.epoch, 0
            self._ncalls[0] += 1
            drop = K.Drop(runtime.rng_state(x.device), (self._site << 32) | (self._ncalls[0] & 0xFFFFFFFF), self.dropout_r)
        x = linear(x, self.linear, self.use_relu, drop)
        if self.dropout_r > 0 and not drop.active:
            x = F.dropout(x, self.dropout_r, self.training)
        return x


class MLP(nn.Module):
    def __init__(self, in_size, mid_size, out_size, dropout_r=0., use_relu=True):
        super().__init__()
        self.fc = FC(in_size, mid_size, dropout_r=dropout_r, use_relu=use_relu)
        self.linear = nn.Linear(mid_size, out_size)

    def forward(self, x):
        return linear(self.fc(x), self.linear)


class AttFlat(nn.Module):
    """Attention pooling head (modules.py:59-85): MLP -> masked softmax over the sequence -> weighted sum -> merge.
    The three dense layers run on the library GEMMs; the softmax over <= 100 positions and the pooling stay torch."""

    def __init__(self, __C):
        super().__init__()
        self.glimpses = __C.ATTFLAT_GLIMPSES
        self.mlp = MLP(__C.HSIZE, __C.ATTFLAT_MLP_SIZE, __C.ATTFLAT_GLIMPSES, dropout_r=__C.DROPOUT_R, use_relu=True)
        self.linear_merge = nn.Linear(__C.HSIZE * __C.ATTFLAT_GLIMPSES, __C.ATTFLAT_OUT_SIZE)

    def forward(self, x, x_mask=None):
        att = self.mlp(x)
        if x_mask is not None:
            att = att.masked_fill(x_mask.squeeze(1).squeeze(1).unsqueeze(2), -1e9)
        att = F.softmax(att, dim=1)
        pooled = torch.einsum('bng,bnh->bgh', att, x).reshape(x.size(0), -1)
        return linear(pooled, self.linear_merge)


class MHAtt(_OpBase):
    """Projection + attention weights of one multi-head attention (bias-free q/k/v/merge)."""
    REL = False

    def __init__(self, __C, base=64, hsize_k=None, bias=False):
        super().__init__()
        self.HBASE = base
        self.HSIZE = __C.HSIZE
        self.HSIZE_INSIDE = int(__C.HSIZE * hsize_k) if hsize_k else __C.HSIZE
        assert self.HSIZE_INSIDE % self.HBASE == 0
        self.HHEAD = self.HSIZE_INSIDE // self.HBASE
        self.DROPOUT_R = __C.DROPOUT_R
        if bias:
            raise NotImplementedError('MHAtt(bias=True) is not used by any registered operator')
        self.linear_v = nn.Linear(__C.HSIZE, self.HSIZE_INSIDE, bias=False)
        self.linear_k = nn.Linear(__C.HSIZE, self.HSIZE_INSIDE, bias=False)
        self.linear_q = nn.Linear(__C.HSIZE, self.HSIZE_INSIDE, bias=False)
        if self.REL:
            self.linear_r = nn.Linear(__C.REL_SIZE, self.HHEAD, bias=True)
        self.linear_merge = nn.Linear(self.HSIZE_INSIDE, __C.HSIZE, bias=False)
        self._init_runtime(2)     # sites: attention map, block output

    def run_block(self, x, kv, mask, rel_embed, ln, residual, out_p):
        """Whole block: LN(x + dropout(self(kv, kv, x)))  — one autograd node, CUDA only."""
        require_cuda(x, kv)
        if self.HBASE != 64:
            raise NotImplementedError("only the '*_64' attention operators (head dim 64) are implemented in CUDA")
        self_att = kv is None or kv is x
        B, Nq = x.shape[0], x.shape[1]
        Nk = Nq if self_att else kv.shape[1]
        mode = self._mode()
        drops = self._drops(x, self.DROPOUT_R)
        d_att = drops[0]
        d_out = drops[1] if out_p > 0 else K.NO_DROP
        w16 = {}
        if mode == 'bf16':
            if self_att:
                w16['vkq'] = self._fused_bf16('vkq', [self.linear_v.weight, self.linear_k.weight, self.linear_q.weight])
            else:
                w16['q'] = self._fused_bf16('q', [self.linear_q.weight])
                w16['vk'] = self._fused_bf16('vk', [self.linear_v.weight, self.linear_k.weight])
            w16['m'] = self._fused_bf16('m', [self.linear_merge.weight])
        cfg = BlockCfg(mode, residual, ln.eps if ln is not None else 1e-6, (d_att, d_out),
                       kmask=_key_mask(mask, B, Nk), x16=_shadow(x), kv16=None if self_att else _shadow(kv), w16=w16)
        rel = g4 = Wy = by = Wr = br = None
        if self.REL:
            assert rel_embed is not None
            Wr, br = self.linear_r.weight, self.linear_r.bias
            if isinstance(rel_embed, RelGeometry):
                g4, Wy, by = rel_embed.g4, rel_embed.weight, rel_embed.bias
            else:
                rel = rel_embed
        fn = AttBlockPyFn if runtime.compose_in_python else AttBlockFn
        out, out16 = fn.apply(x, None if self_att else kv, self.linear_q.weight, self.linear_k.weight,
                                      self.linear_v.weight, self.linear_merge.weight,
                                      ln.a_2 if ln is not None else None, ln.b_2 if ln is not None else None,
                                      rel, g4, Wy, by, Wr, br, cfg)
        if out16 is not None:
            out._mmnas_bf16 = out16
        return out

    def forward(self, v, k, q, mask=None, rel_embed=None):
        if v is not k:
            raise NotImplementedError('values and keys must come from the same tensor (as in every registered operator)')
        return self.run_block(q, None if k is q else k, mask, rel_embed, None, False, 0.0)


class RelMHAtt(MHAtt):
    REL = True

    def forward(self, v, k, q, mask=None, rel_embed=None):
        assert rel_embed is not None
        return super().forward(v, k, q, mask, rel_embed)


class _AttOp(nn.Module):
    MH = MHAtt

    def __init__(self, __C, norm=False, residual=False, base=64, hsize_k=None):
        super().__init__()
        self.norm = norm
        self.residual = residual
        self.DROPOUT_R = __C.DROPOUT_R
        self.mhatt = self.MH(__C, base=base, hsize_k=hsize_k)
        self.dropout = nn.Dropout(__C.DROPOUT_R)   # kept for module-tree parity; applied inside the fused kernel
        if norm:
            self.ln = LayerNorm(__C.HSIZE)

    GUIDED = False

    def shadow_specs(self):
        """(owner module, shadow name, [weights stacked along dim 0]) for engine.WeightShadows."""
        m = self.mhatt
        if self.GUIDED:
            specs = [(m, 'q', [m.linear_q.weight]), (m, 'vk', [m.linear_v.weight, m.linear_k.weight])]
        else:
            specs = [(m, 'vkq', [m.linear_v.weight, m.linear_k.weight, m.linear_q.weight])]
        return specs + [(m, 'm', [m.linear_merge.weight])]

    def _run(self, x, kv, mask, rel_embed):
        return self.mhatt.run_block(x, kv, mask, rel_embed, self.ln if self.norm else None, self.residual,
                                    self.DROPOUT_R)


class SelfAtt(_AttOp):
    def forward(self, x, y=None, x_mask=None, y_mask=None, rel_embed=None):
        return self._run(x, None, x_mask, None)


class RelSelfAtt(_AttOp):
    MH = RelMHAtt

    def forward(self, x, y=None, x_mask=None, y_mask=None, rel_embed=None):
        assert rel_embed is not None
        return self._run(x, None, x_mask, rel_embed)


class GuidedAtt(_AttOp):
    GUIDED = True

    def forward(self, x, y=None, x_mask=None, y_mask=None, rel_embed=None):
        assert y is not None
        return self._run(x, y, y_mask, None)


class FeedForward(_OpBase):
    def __init__(self, __C, norm=False, residual=False, mid_k=None):
        super().__init__()
        self.norm = norm
        self.residual = residual
        self.DROPOUT_R = __C.DROPOUT_R
        self.MID_SIZE = __C.HSIZE * (mid_k if mid_k else 4)
        self.mlp = MLP(__C.HSIZE, self.MID_SIZE, __C.HSIZE, dropout_r=__C.DROPOUT_R, use_relu=True)
        self.dropout = nn.Dropout(__C.DROPOUT_R)
        if norm:
            self.ln = LayerNorm(__C.HSIZE)
        self._init_runtime(2)     # sites: hidden activation, block output

    def shadow_specs(self):
        return [(self, 'w1', [self.mlp.fc.linear.weight]), (self, 'w2', [self.mlp.linear.weight])]

    def forward(self, x, y=None, x_mask=None, y_mask=None, rel_embed=None):
        require_cuda(x)
        mode = self._mode()
        w1, w2 = self.mlp.fc.linear, self.mlp.linear
        w16 = {}
        if mode == 'bf16':
            w16['w1'] = self._fused_bf16('w1', [w1.weight])
            w16['w2'] = self._fused_bf16('w2', [w2.weight])
        ln = self.ln if self.norm else None
        cfg = BlockCfg(mode, self.residual, ln.eps if ln is not None else 1e-6, self._drops(x, self.DROPOUT_R),
                       x16=_shadow(x), w16=w16)
        fn = FFNBlockPyFn if runtime.compose_in_python else FFNBlockFn
        out, out16 = fn.apply(x, w1.weight, w1.bias, w2.weight, w2.bias,
                                      ln.a_2 if ln is not None else None, ln.b_2 if ln is not None else None, cfg)
        if out16 is not None:
            out._mmnas_bf16 = out16
        return out


class Identity(nn.Module):
    def forward(self, x, y=None, x_mask=None, y_mask=None, rel_embed=None):
        return x


class Zero(nn.Module):
    def forward(self, x, y=None, x_mask=None, y_mask=None, rel_embed=None):
        return x * 0.
