"""Block-level autograd Functions of the operator hot path.

Each candidate block (SA / GA / RSA / FFN, reference mmnas/model/modules.py:248-362) is ONE autograd node whose
forward and backward are fixed sequences of C-ABI kernel calls (mmnas_b200.kernels) — no torch arithmetic.
Two arms:
  fp32  FFMA GEMMs + fp32 attention                      (parity gate 1e-5 normwise)
  bf16  tcgen05 GEMMs on bf16 operands, fp32 accumulate,  (parity gate 2e-2 normwise)
        fp32 softmax / LayerNorm / geometry statistics
The residual stream (block inputs/outputs) is fp32 in both arms; the bf16 arm additionally emits a bf16
shadow of each block output so the next block's GEMM reads it without a cast pass.
"""
import math

import torch
from torch.autograd import Function

from . import _lib
from . import kernels as K
from . import runtime
from ._lib import require_cuda

HEAD = 64


class BlockCfg:
    """Non-tensor arguments of one block call."""

    def __init__(self, precision, residual, eps, drops, kmask=None, x16=None, kv16=None, w16=None):
        self.precision = precision      # 'fp32' | 'bf16'
        self.residual = residual
        self.eps = eps
        self.drops = drops              # tuple of kernels.Drop, one per dropout site of the block
        self.kmask = kmask              # uint8 [B, Nk] (1 = padded key) or None
        self.x16, self.kv16 = x16, kv16 # optional bf16 shadows of the inputs
        self.w16 = w16 or {}            # bf16 (fused) weight copies for the tensor-core arm


def _split_k(M, N, K):
    """Split the token reduction of a weight-gradient GEMM so that (#output tiles x splits) ~ one wave of the GPU,
    keeping at least 4 k-blocks of 64 per split.  Large outputs with a long reduction go to the CTA-pair kernel
    (256 x 256 tiles, 74 clusters; the library switches to it at >= 48 units), everything else to 128 x 128 tiles on
    148 SMs — measured in scripts/bench_gemm_pair.py."""
    kb = (K + 63) // 64
    cap = kb // 4 if kb >= 4 else 1
    if N % 256 == 0 and M >= 256 and K >= 2048:
        units = ((M + 255) // 256) * (N // 256)
        if 8 <= units <= 74:
            return max(1, min(74 // units, cap))
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    return max(1, min(148 // tiles if tiles <= 148 else 1, cap))


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


def _grad_target(params):
    """When the engine asked for direct gradient accumulation (runtime.direct_grads) and `params` own preset,
    contiguous fp32 .grad buffers that sit back to back in memory (engine.FlatGrads), return one fp32 view
    [sum rows, cols] over those buffers; the backward then accumulates into it (split-K red.add / C += ...) and returns
    None for these parameters, which removes autograd's per-parameter `grad += new` kernels and the zero-filled
    temporaries.  Otherwise None."""
    if not runtime.direct_grads:
        return None
    if any(getattr(p, '_mmnas_shared', False) for p in params):
        return None     # used by several blocks per step: autograd must sum the contributions and fire its hook ONCE
    g0 = params[0].grad
    if g0 is None:
        return None
    rows, ptr = 0, g0.data_ptr()
    for p in params:
        g = p.grad
        if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() != ptr:
            return None
        ptr += 4 * g.numel()
        rows += p.shape[0]
    shape = (rows,) + tuple(params[0].shape[1:])
    return torch.as_strided(g0, shape, tuple(params[0].stride()) if params[0].dim() > 1 else (1,))


import os as _os
_FORK_CTX = _os.environ.get('MMNAS_FORK_CTX', '0') == '1'     # A/B only: torch.cuda.stream() context instead of the launch-stream override


class _Fork:
    """Weight-gradient work of one block backward on the side stream: `with fork:` enqueues there after everything
    issued so far on the main stream; join() makes the main stream wait for it (called before the backward returns,
    so every buffer the side work touches is still referenced and later reuse of its memory is ordered after it).
    Captured in a CUDA graph this becomes a parallel branch; SMs left idle by a 100-CTA dgrad GEMM run wgrad CTAs.

    INVARIANT: inside `with fork:` only the library's launches move to the side stream — torch's current stream and the
    caching allocator still see the main stream.  No torch kernel may be issued there: a `torch.zeros` destination
    filled on the main stream races the side-stream kernel that accumulates into it.  Prepare every destination
    (`_Sink.prepare`, temporaries) BEFORE entering the block; `_Sink.prepare` raises if it is called inside one."""
    depth = 0

    def __init__(self, dev, enabled):
        self.enabled = enabled and runtime.overlap_wgrad
        self.used = False
        if self.enabled:
            self.main = torch.cuda.current_stream(dev)
            self.side = runtime.side_stream(dev)

    def __enter__(self):       # only library kernels are launched inside the block: redirect them, not torch's stream
        if self.enabled:
            self.side.wait_stream(self.main)
            if _FORK_CTX:
                self.ctx = torch.cuda.stream(self.side)
                self.ctx.__enter__()
            else:
                _lib.set_stream_override(self.side.cuda_stream)
            self.used = True
            _Fork.depth += 1
        return self

    def __exit__(self, *a):
        if self.enabled:
            _Fork.depth -= 1
            if _FORK_CTX:
                self.ctx.__exit__(*a)
            else:
                _lib.set_stream_override(None)
        return False

    def join(self):
        if self.enabled and self.used:
            self.main.wait_stream(self.side)


class _Sink:
    """Destination of a parameter gradient: the parameters' own .grad memory (direct mode: the kernel accumulates,
    autograd gets None) or a fresh buffer returned to autograd."""

    def __init__(self, params, dev):
        self.params = params
        self.dev = dev
        self.target = _grad_target(params)
        self.direct = self.target is not None
        self.buf = self.target

    def prepare(self, zero):
        if _Fork.depth and not _FORK_CTX:
            raise RuntimeError('_Sink.prepare() inside `with fork:`: the fill would run on the main stream and race the '
                               'side-stream kernels (see _Fork)')
        if not self.direct:
            rows = sum(p.shape[0] for p in self.params)
            shape = (rows,) + tuple(self.params[0].shape[1:])
            self.buf = (torch.zeros if zero else torch.empty)(shape, dtype=torch.float32, device=self.dev)
        return self

    def grads(self):
        """Call after fork.join(): in direct mode this tells the data-parallel reducer the gradients are complete,
        and the reducer may enqueue the bucket's all-reduce behind the CURRENT stream right away."""
        if self.direct:
            runtime.notify_grads(self.params)
            return (None,) * len(self.params)
        out, r = [], 0
        for p in self.params:
            out.append(self.buf[r:r + p.shape[0]])
            r += p.shape[0]
        return tuple(out)


def _bf16(t, shadow):
    if shadow is not None:
        return shadow
    return K.cast_bf16(t)


# ----------------------------------------------------------------------------------------------------------
# attention blocks: SelfAtt / RelSelfAtt / GuidedAtt  (modules.py:248-325 over MHAtt :158-199, RelMHAtt :202-245)
# ----------------------------------------------------------------------------------------------------------
class AttBlockPyFn(Function):
    """out = LN(x + dropout(merge(att(q(x), k(kv), v(kv)))));  kv=None means self-attention (kv is x).
    The block composed from the PRIMITIVE entry points, one foreign call per kernel (runtime.compose_in_python): the
    decomposed path bench.py times kernel by kernel, and the cross-check of the C-composed block (AttBlockFn)."""

    @staticmethod
    def forward(ctx, x, kv, Wq, Wk, Wv, Wm, a2, b2, rel, g4, Wy, by, Wr, br, cfg):
        require_cuda(x, kv, Wq)
        ctx.set_materialize_grads(False)      # no zero-filled gradient tensor for the bf16 shadow output
        dev = x.device
        x = x.contiguous()
        B, Nq, H = x.shape
        I = Wq.shape[0]
        heads = I // HEAD
        self_att = kv is None
        kvt = x if self_att else kv.contiguous()
        Nk = kvt.shape[1]
        Mq, Mk = B * Nq, B * Nk
        bf = cfg.precision == 'bf16'
        adt = torch.bfloat16 if bf else torch.float32
        d_att, d_out = cfg.drops
        scale = 1.0 / math.sqrt(HEAD)

        x16 = kv16 = None
        if bf:
            x16 = _bf16(x, cfg.x16)
            kv16 = x16 if self_att else _bf16(kvt, cfg.kv16)
        # --- projections
        if self_att:
            qkv = _empty((Mq, 3 * I), adt, dev)          # columns [v | k | q]: the parameters' registration order
            v, k, q = qkv[:, :I], qkv[:, I:2 * I], qkv[:, 2 * I:]
            kvb = None
            if bf:
                K.gemm_bf16(Mq, 3 * I, H, x16, H, 0, cfg.w16['vkq'], H, 0, qkv, 3 * I)
            else:
                for W, dst in ((Wq, q), (Wk, k), (Wv, v)):
                    K.gemm_f32(Mq, I, H, x, H, 1, W, 1, H, dst, 3 * I)
        else:
            qkv = _empty((Mq, I), adt, dev)
            kvb = _empty((Mk, 2 * I), adt, dev)          # columns [v | k]
            q, v, k = qkv, kvb[:, :I], kvb[:, I:]
            if bf:
                K.gemm_bf16(Mq, I, H, x16, H, 0, cfg.w16['q'], H, 0, qkv, I)
                K.gemm_bf16(Mk, 2 * I, H, kv16, H, 0, cfg.w16['vk'], H, 0, kvb, 2 * I)
            else:
                K.gemm_f32(Mq, I, H, x, H, 1, Wq, 1, H, q, I)
                K.gemm_f32(Mk, I, H, kvt, H, 1, Wk, 1, H, k, 2 * I)
                K.gemm_f32(Mk, I, H, kvt, H, 1, Wv, 1, H, v, 2 * I)
        # --- RSA logit bias from the geometry path (fp32 kernels; tensor-core arithmetic with split operands in the bf16 arm)
        bias = None
        R = 0
        if Wr is not None:
            R = Wr.shape[1]
            bias = _empty((B, heads, Nq, Nk), torch.float32, dev)
            if g4 is not None:
                K.relbias_fwd(B, Nq, heads, R, None, g4.contiguous(), Wy, by, Wr, br, bias, mode=1 if bf else 0)
            else:
                rel = rel.contiguous()
                K.relbias_fwd(B, Nq, heads, R, rel, None, None, None, Wr, br, bias)
        # --- attention core
        atted = _empty((Mq, I), adt, dev)
        K.attn_fwd(B, heads, Nq, Nk, q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                   cfg.kmask, bias, atted, I, scale, d_att)
        # --- merge projection, residual, LayerNorm
        branch = _empty((Mq, H), torch.float32, dev)
        if bf:
            K.gemm_bf16(Mq, H, I, atted, I, 0, cfg.w16['m'], I, 0, branch, H)
        else:
            K.gemm_f32(Mq, H, I, atted, I, 1, Wm, 1, I, branch, H)
        out = _empty((B, Nq, H), torch.float32, dev)
        out16 = _empty((B, Nq, H), torch.bfloat16, dev) if bf else None
        norm = a2 is not None
        mean = _empty((Mq,), torch.float32, dev) if norm else None
        sigma = _empty((Mq,), torch.float32, dev) if norm else None
        K.ln_residual_fwd(Mq, H, x if cfg.residual else None, branch, a2, b2, cfg.eps, out, out16, mean, sigma, d_out)

        ctx.cfg, ctx.dims = cfg, (B, Nq, Nk, H, I, heads, R, self_att)
        ctx.b2 = b2
        ctx.save_for_backward(x, kvt, Wq, Wk, Wv, Wm, a2, rel, g4, Wy, by, Wr, br, qkv, kvb, bias, atted, branch, mean,
                              sigma, x16, kv16)
        if bf:
            ctx.mark_non_differentiable(out16)
            return out, out16
        return out, None

    @staticmethod
    def backward(ctx, dout, _unused=None):
        if dout is None:
            return (None,) * 15
        cfg = ctx.cfg
        B, Nq, Nk, H, I, heads, R, self_att = ctx.dims
        (x, kvt, Wq, Wk, Wv, Wm, a2, rel, g4, Wy, by, Wr, br, qkv, kvb, bias, atted, z, mean, sigma, x16,
         kv16) = ctx.saved_tensors
        dev = x.device
        Mq, Mk = B * Nq, B * Nk
        bf = cfg.precision == 'bf16'
        adt = torch.bfloat16 if bf else torch.float32
        d_att, d_out = cfg.drops
        scale = 1.0 / math.sqrt(HEAD)
        norm = a2 is not None
        dout = dout.contiguous()
        fork = _Fork(dev, bf)

        def wgrad(sink, M_, N_, K_, A, lda, Bm, ldb, a32, b32):
            """sink (+)= A^T Bm over K_ tokens.  bf16: operands [K_, M_] / [K_, N_] read MN-major; fp32: strided FFMA.
            The sink buffer is allocated (and zeroed) on the main stream, only the GEMM may run on the side stream."""
            if bf:
                sk = _split_k(M_, N_, K_)
                sink.prepare(zero=sk > 1)
                with fork:
                    K.gemm_bf16(M_, N_, K_, A, lda, 1, Bm, ldb, 1, sink.buf, N_, split_k=sk,
                                accumulate=sink.direct and sk == 1)
            else:
                sink.prepare(zero=False)
                K.gemm_f32(M_, N_, K_, a32, 1, lda, b32, ldb, 1, sink.buf, N_, accumulate=sink.direct)

        # --- LayerNorm + residual + output-dropout backward
        dz = _empty((Mq, H), torch.float32, dev) if cfg.residual else None
        separate = bf or d_out.active or not cfg.residual
        dbranch = _empty((Mq, H), adt, dev) if separate else None
        s_a2 = s_b2 = None
        if norm:
            s_a2, s_b2 = _Sink([a2], dev).prepare(True), _Sink([ctx.b2], dev).prepare(True)
        K.ln_residual_bwd(Mq, H, dout, z, mean, sigma, a2, cfg.eps, dz, dbranch, s_a2.buf if norm else None,
                          s_b2.buf if norm else None, d_out)
        if dbranch is None:
            dbranch = dz
        # --- merge projection backward
        datt = _empty((Mq, I), adt, dev)
        s_m = _Sink([Wm], dev)
        wgrad(s_m, H, I, Mq, dbranch, H, atted, I, dbranch, atted)
        if bf:
            K.gemm_bf16(Mq, I, H, dbranch, H, 0, cfg.w16['m'], I, 1, datt, I)
        else:
            K.gemm_f32(Mq, I, H, dbranch, H, 1, Wm, I, 1, datt, I)
        # --- attention core backward (fused buffers keep the parameters' registration order v, k, q)
        if self_att:
            v, k, q = qkv[:, :I], qkv[:, I:2 * I], qkv[:, 2 * I:]
            dqkv = _empty((Mq, 3 * I), adt, dev)
            dv, dk, dq = dqkv[:, :I], dqkv[:, I:2 * I], dqkv[:, 2 * I:]
            dkvb = None
        else:
            q, v, k = qkv, kvb[:, :I], kvb[:, I:]
            dqkv = _empty((Mq, I), adt, dev)
            dkvb = _empty((Mk, 2 * I), adt, dev)
            dq, dv, dk = dqkv, dkvb[:, :I], dkvb[:, I:]
        dbias = _empty((B, heads, Nq, Nk), torch.float32, dev) if bias is not None else None
        K.attn_bwd(B, heads, Nq, Nk, q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                   cfg.kmask, bias, atted, I, datt, I, dq.data_ptr(), dq.stride(0), dk.data_ptr(), dk.stride(0),
                   dv.data_ptr(), dv.stride(0), dbias, scale, d_att)
        # --- geometry-bias backward (kernels accumulate with atomics)
        drel = None
        g_Wy = g_by = g_Wr = g_br = None
        if bias is not None:
            s_Wr, s_br = _Sink([Wr], dev).prepare(True), _Sink([br], dev).prepare(True)
            if g4 is not None:
                s_Wy, s_by = _Sink([Wy], dev).prepare(True), _Sink([by], dev).prepare(True)
                K.relbias_bwd(B, Nq, heads, R, None, g4, Wy, by, Wr, br, dbias, None, s_Wy.buf, s_by.buf, s_Wr.buf, s_br.buf,
                              mode=1 if bf else 0)
                g_Wy, g_by = s_Wy.grads()[0], s_by.grads()[0]
            else:
                drel = torch.empty_like(rel)
                K.relbias_bwd(B, Nq, heads, R, rel, None, None, None, Wr, br, dbias, drel, None, None, s_Wr.buf, s_br.buf)
            g_Wr, g_br = s_Wr.grads()[0], s_br.grads()[0]
        # --- projection backward: weight gradients and input gradients
        dkv_in = None
        fresh_dz = dz is None
        if fresh_dz:
            dz = _empty((Mq, H), torch.float32, dev)
        if self_att:
            s_vkq = _Sink([Wv, Wk, Wq], dev)
            wgrad(s_vkq, 3 * I, H, Mq, dqkv, 3 * I, x16, H, dqkv, x)
            if bf:
                K.gemm_bf16(Mq, H, 3 * I, dqkv, 3 * I, 0, cfg.w16['vkq'], H, 1, dz, H, accumulate=not fresh_dz)
            else:
                acc = not fresh_dz
                for W, d in ((Wv, dv), (Wk, dk), (Wq, dq)):
                    K.gemm_f32(Mq, H, I, d, 3 * I, 1, W, H, 1, dz, H, accumulate=acc)
                    acc = True
        else:
            dkv_in = _empty((Mk, H), torch.float32, dev)
            s_q, s_vk = _Sink([Wq], dev), _Sink([Wv, Wk], dev)
            wgrad(s_q, I, H, Mq, dqkv, I, x16, H, dqkv, x)
            wgrad(s_vk, 2 * I, H, Mk, dkvb, 2 * I, kv16, H, dkvb, kvt)
            if bf:
                K.gemm_bf16(Mq, H, I, dqkv, I, 0, cfg.w16['q'], H, 1, dz, H, accumulate=not fresh_dz)
                K.gemm_bf16(Mk, H, 2 * I, dkvb, 2 * I, 0, cfg.w16['vk'], H, 1, dkv_in, H)
            else:
                K.gemm_f32(Mq, H, I, dq, I, 1, Wq, H, 1, dz, H, accumulate=not fresh_dz)
                K.gemm_f32(Mk, H, I, dv, 2 * I, 1, Wv, H, 1, dkv_in, H)
                K.gemm_f32(Mk, H, I, dk, 2 * I, 1, Wk, H, 1, dkv_in, H, accumulate=True)
            dkv_in = dkv_in.view(B, Nk, H)
        fork.join()
        if self_att:
            g_Wv, g_Wk, g_Wq = s_vkq.grads()
        else:
            g_Wq, = s_q.grads()
            g_Wv, g_Wk = s_vk.grads()
        g_a2 = s_a2.grads()[0] if norm else None
        g_b2 = s_b2.grads()[0] if norm else None
        return (dz.view(B, Nq, H), dkv_in, g_Wq, g_Wk, g_Wv, s_m.grads()[0], g_a2, g_b2, drel, None, g_Wy, g_by, g_Wr,
                g_br, None)


# ----------------------------------------------------------------------------------------------------------
# FeedForward  (modules.py:328-362 over MLP :34-41 / FC :13-31)
# ----------------------------------------------------------------------------------------------------------
class FFNBlockPyFn(Function):
    """out = LN(x + dropout(W2 dropout(relu(W1 x + b1)) + b2)) composed from the primitive entry points (see AttBlockPyFn)."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2f, a2, b2, cfg):
        require_cuda(x, W1)
        ctx.set_materialize_grads(False)
        dev = x.device
        x = x.contiguous()
        B, N, H = x.shape
        Fd = W1.shape[0]
        M = B * N
        bf = cfg.precision == 'bf16'
        adt = torch.bfloat16 if bf else torch.float32
        d_mid, d_out = cfg.drops
        x16 = _bf16(x, cfg.x16) if bf else None
        h = _empty((M, Fd), adt, dev)
        branch = _empty((M, H), torch.float32, dev)
        if bf:
            K.gemm_bf16(M, Fd, H, x16, H, 0, cfg.w16['w1'], H, 0, h, Fd, bias=b1, relu=True, drop=d_mid)
            K.gemm_bf16(M, H, Fd, h, Fd, 0, cfg.w16['w2'], Fd, 0, branch, H, bias=b2f)
        else:
            K.gemm_f32(M, Fd, H, x, H, 1, W1, 1, H, h, Fd, bias=b1, epilogue=2 if d_mid.active else 1, drop=d_mid)
            K.gemm_f32(M, H, Fd, h, Fd, 1, W2, 1, Fd, branch, H, bias=b2f)
        out = _empty((B, N, H), torch.float32, dev)
        out16 = _empty((B, N, H), torch.bfloat16, dev) if bf else None
        norm = a2 is not None
        mean = _empty((M,), torch.float32, dev) if norm else None
        sigma = _empty((M,), torch.float32, dev) if norm else None
        K.ln_residual_fwd(M, H, x if cfg.residual else None, branch, a2, b2, cfg.eps, out, out16, mean, sigma, d_out)
        ctx.cfg, ctx.dims = cfg, (B, N, H, Fd)
        ctx.small = (b1, b2f, b2)
        ctx.save_for_backward(x, W1, W2, a2, h, branch, mean, sigma, x16)
        if bf:
            ctx.mark_non_differentiable(out16)
            return out, out16
        return out, None

    @staticmethod
    def backward(ctx, dout, _unused=None):
        if dout is None:
            return (None,) * 8
        cfg = ctx.cfg
        B, N, H, Fd = ctx.dims
        x, W1, W2, a2, h, z, mean, sigma, x16 = ctx.saved_tensors
        b1, b2f, b2 = ctx.small
        dev = x.device
        M = B * N
        bf = cfg.precision == 'bf16'
        adt = torch.bfloat16 if bf else torch.float32
        d_mid, d_out = cfg.drops
        norm = a2 is not None
        dout = dout.contiguous()
        dz = _empty((M, H), torch.float32, dev) if cfg.residual else None
        separate = bf or d_out.active or not cfg.residual
        dbranch = _empty((M, H), adt, dev) if separate else None
        s_a2 = s_b2 = None
        if norm:
            s_a2, s_b2 = _Sink([a2], dev).prepare(True), _Sink([b2], dev).prepare(True)
        K.ln_residual_bwd(M, H, dout, z, mean, sigma, a2, cfg.eps, dz, dbranch, s_a2.buf if norm else None,
                          s_b2.buf if norm else None, d_out)
        if dbranch is None:
            dbranch = dz
        keep_scale = 1.0 / (1.0 - d_mid.p) if d_mid.active else 1.0
        s_W1, s_W2, s_bias1, s_bias2 = _Sink([W1], dev), _Sink([W2], dev), _Sink([b1], dev), _Sink([b2f], dev)
        dh = _empty((M, Fd), adt, dev)
        fork = _Fork(dev, bf)
        fresh_dz = dz is None
        if fresh_dz:
            dz = _empty((M, H), torch.float32, dev)
        if bf:
            sk2, sk1 = _split_k(H, Fd, M), _split_k(Fd, H, M)
            s_bias2.prepare(False)
            s_W2.prepare(zero=sk2 > 1)
            s_bias1.prepare(False)
            s_W1.prepare(zero=sk1 > 1)
            with fork:       # db2, dW2 need only dbranch: they overlap the dh GEMM
                K.colsum(dbranch, M, H, H, s_bias2.buf, accumulate=s_bias2.direct)
                K.gemm_bf16(H, Fd, M, dbranch, H, 1, h, Fd, 1, s_W2.buf, Fd, split_k=sk2, accumulate=s_W2.direct and sk2 == 1)
            K.gemm_bf16(M, Fd, H, dbranch, H, 0, cfg.w16['w2'], Fd, 1, dh, Fd, aux=h, ld_aux=Fd, aux_scale=keep_scale)
            with fork:       # db1, dW1 need dh: they overlap the dx GEMM
                K.colsum(dh, M, Fd, Fd, s_bias1.buf, accumulate=s_bias1.direct)
                K.gemm_bf16(Fd, H, M, dh, Fd, 1, x16, H, 1, s_W1.buf, H, split_k=sk1, accumulate=s_W1.direct and sk1 == 1)
            K.gemm_bf16(M, H, Fd, dh, Fd, 0, cfg.w16['w1'], H, 1, dz, H, accumulate=not fresh_dz)
        else:
            s_bias2.prepare(False)
            K.colsum(dbranch, M, H, H, s_bias2.buf, accumulate=s_bias2.direct)
            s_W2.prepare(False)
            K.gemm_f32(H, Fd, M, dbranch, 1, H, h, Fd, 1, s_W2.buf, Fd, accumulate=s_W2.direct)
            K.gemm_f32(M, Fd, H, dbranch, H, 1, W2, Fd, 1, dh, Fd, epilogue=3, aux=h, ld_aux=Fd, aux_scale=keep_scale)
            s_bias1.prepare(False)
            K.colsum(dh, M, Fd, Fd, s_bias1.buf, accumulate=s_bias1.direct)
            s_W1.prepare(False)
            K.gemm_f32(Fd, H, M, dh, 1, Fd, x, H, 1, s_W1.buf, H, accumulate=s_W1.direct)
            K.gemm_f32(M, H, Fd, dh, Fd, 1, W1, H, 1, dz, H, accumulate=not fresh_dz)
        fork.join()
        return (dz.view(B, N, H), s_W1.grads()[0], s_bias1.grads()[0], s_W2.grads()[0], s_bias2.grads()[0],
                s_a2.grads()[0] if norm else None, s_b2.grads()[0] if norm else None, None)


# ----------------------------------------------------------------------------------------------------------
# Block-level calls (ABI v7): one foreign call per block forward / backward.  The C side (csrc/blocks.cu) carves
# the workspace, orders the launches and forks the weight-gradient GEMMs onto the side stream.
# ----------------------------------------------------------------------------------------------------------
def _ptr(t):
    return None if t is None else t.data_ptr()


def _direct(params):
    """Gradient destinations for direct accumulation: every parameter owns a preset contiguous fp32 .grad (a view of
    engine.FlatGrads) and none is shared between blocks.  Returns the .grad tensors or None."""
    if not runtime.direct_grads:
        return None
    out = []
    for p in params:
        if p is None:
            out.append(None)
            continue
        g = p.grad
        if g is None or g.dtype != torch.float32 or not g.is_contiguous() or getattr(p, '_mmnas_shared', False):
            return None
        out.append(g)
    return out


def _side_stream(dev, bf):
    if bf and runtime.overlap_wgrad:
        return runtime.side_stream(dev).cuda_stream
    return None


class AttBlockFn(Function):
    """SelfAtt / GuidedAtt / RelSelfAtt as ONE foreign call per direction (mmnas_mha_ln_* / mmnas_rel_mha_ln_*)."""

    @staticmethod
    def forward(ctx, x, kv, Wq, Wk, Wv, Wm, a2, b2, rel, g4, Wy, by, Wr, br, cfg):
        require_cuda(x, kv, Wq)
        ctx.set_materialize_grads(False)
        dev = x.device
        x = x.contiguous()
        B, Nq, H = x.shape
        I = Wq.shape[0]
        guided = kv is not None
        kvt = kv.contiguous() if guided else None
        Nk = kvt.shape[1] if guided else Nq
        bf = cfg.precision == 'bf16'
        d_att, d_out = cfg.drops
        d = _lib.AttBlock()
        d.precision = 1 if bf else 0
        d.B, d.Nq, d.Nk, d.H, d.I = B, Nq, Nk, H, I
        d.R = Wr.shape[1] if Wr is not None else 0
        d.residual = 1 if cfg.residual else 0
        d.guided = 1 if guided else 0
        d.eps = cfg.eps
        st = d_att.state if d_att.active else (d_out.state if d_out.active else None)
        if st is not None:
            d.rng_state = st.data_ptr()
            if d_att.active:
                d.p_att, d.salt_att = d_att.p, d_att.salt
            if d_out.active:
                d.p_out, d.salt_out = d_out.p, d_out.salt
        x16 = kv16 = None
        if bf:
            x16 = cfg.x16
            kv16 = cfg.kv16 if guided else None
            w = cfg.w16
            d.w16_a = (w['q'] if guided else w['vkq']).data_ptr()
            if guided:
                d.w16_b = w['vk'].data_ptr()
            d.w16_m = w['m'].data_ptr()
        d.x, d.x16 = x.data_ptr(), _ptr(x16)
        if guided:
            d.kv, d.kv16 = kvt.data_ptr(), _ptr(kv16)
        d.kmask = _ptr(cfg.kmask)
        d.Wq, d.Wk, d.Wv, d.Wm = Wq.data_ptr(), Wk.data_ptr(), Wv.data_ptr(), Wm.data_ptr()
        d.ln_a, d.ln_b = _ptr(a2), _ptr(b2)
        if d.R:
            if g4 is not None:
                g4 = g4.contiguous()
                d.g4, d.Wy, d.by = g4.data_ptr(), Wy.data_ptr(), by.data_ptr()
            else:
                rel = rel.contiguous()
                d.rel = rel.data_ptr()
            d.Wr, d.br = Wr.data_ptr(), br.data_ptr()
        fwd_bytes, bwd_bytes = _lib.workspace_bytes(d)
        ws = torch.empty(fwd_bytes, dtype=torch.uint8, device=dev)
        out = torch.empty((B, Nq, H), dtype=torch.float32, device=dev)
        out16 = torch.empty((B, Nq, H), dtype=torch.bfloat16, device=dev) if bf else None
        d.out, d.out16, d.workspace = out.data_ptr(), _ptr(out16), ws.data_ptr()
        d.stream = _lib.stream()
        _lib.call_block('mmnas_rel_mha_ln_fwd' if d.R else 'mmnas_mha_ln_fwd', d)
        ctx.desc, ctx.bwd_bytes, ctx.guided, ctx.bf = d, bwd_bytes, guided, bf
        ctx.params = (Wq, Wk, Wv, Wm, a2, b2, Wr, br)
        ctx.geo = (Wy, by)
        ctx.save_for_backward(x, kvt, rel, g4, ws, x16, kv16, cfg.kmask)
        ctx.w16 = cfg.w16            # keeps the bf16 weight copies alive until the backward has run
        if bf:
            ctx.mark_non_differentiable(out16)
            return out, out16
        return out, None

    @staticmethod
    def backward(ctx, dout, _unused=None):
        if dout is None:
            return (None,) * 15
        d = ctx.desc
        x, kvt, rel, g4, ws, x16, kv16, kmask = ctx.saved_tensors
        Wq, Wk, Wv, Wm, a2, b2, Wr, br = ctx.params
        Wy, by = ctx.geo
        dev = x.device
        B, Nq, Nk, H, I = d.B, d.Nq, d.Nk, d.H, d.I
        dout = dout.contiguous()
        dx = torch.empty((B, Nq, H), dtype=torch.float32, device=dev)
        dkv = torch.empty((B, Nk, H), dtype=torch.float32, device=dev) if ctx.guided else None
        bws = torch.empty(ctx.bwd_bytes, dtype=torch.uint8, device=dev)
        own = (Wv, Wk, Wq, Wm, a2, b2, Wr, br)
        sinks = _direct(own)
        direct = sinks is not None
        if not direct:
            f32 = dict(dtype=torch.float32, device=dev)
            if ctx.guided:
                gq, gvk = torch.empty((I, H), **f32), torch.empty((2 * I, H), **f32)
                gv, gk = gvk[:I], gvk[I:]
            else:
                gvkq = torch.empty((3 * I, H), **f32)
                gv, gk, gq = gvkq[:I], gvkq[I:2 * I], gvkq[2 * I:]
            gm = torch.empty((H, I), **f32)
            ga = gb = gr = gbr = None
            if a2 is not None:
                gab = torch.empty((2, H), **f32)
                ga, gb = gab[0], gab[1]
            if d.R:
                grb = torch.empty((Wr.shape[0] * (d.R + 1),), **f32)
                gr, gbr = grb[:Wr.numel()].view_as(Wr), grb[Wr.numel():]
            sinks = [gv, gk, gq, gm, ga, gb, gr, gbr]
        d.accumulate_grads = 1 if direct else 0
        d.dWv, d.dWk, d.dWq, d.dWm = (sinks[i].data_ptr() for i in range(4))
        d.dln_a, d.dln_b, d.dWr, d.dbr = _ptr(sinks[4]), _ptr(sinks[5]), _ptr(sinks[6]), _ptr(sinks[7])
        gy = gby = drel = None
        if d.R:
            if g4 is not None:
                geo = _direct((Wy, by))
                if geo is not None:
                    d.accumulate_geometry, gyb = 1, None
                    d.dWy, d.dby = geo[0].data_ptr(), geo[1].data_ptr()
                else:                   # linear_y_rel is shared by every RSA block: autograd sums the contributions
                    d.accumulate_geometry = 0
                    gyb = torch.empty((d.R * 5,), dtype=torch.float32, device=dev)
                    gy, gby = gyb[:d.R * 4].view(d.R, 4), gyb[d.R * 4:]
                    d.dWy, d.dby = gy.data_ptr(), gby.data_ptr()
            else:
                drel = torch.empty_like(rel)
                d.drel = drel.data_ptr()
        d.dout, d.dx, d.dkv, d.bwd_workspace = dout.data_ptr(), dx.data_ptr(), _ptr(dkv), bws.data_ptr()
        d.stream = _lib.stream()
        d.side_stream = _side_stream(dev, ctx.bf)
        _lib.call_block('mmnas_rel_mha_ln_bwd' if d.R else 'mmnas_mha_ln_bwd', d)
        if direct:
            runtime.notify_grads(own)
            if d.R and g4 is not None and d.accumulate_geometry:
                runtime.notify_grads((Wy, by))
            gv = gk = gq = gm = ga = gb = gr = gbr = None
        else:
            gv, gk, gq, gm, ga, gb, gr, gbr = sinks
        return (dx, dkv, gq, gk, gv, gm, ga, gb, drel, None, gy, gby, gr, gbr, None)


class FFNBlockFn(Function):
    """FeedForward as ONE foreign call per direction (mmnas_ffn_ln_fwd / mmnas_ffn_ln_bwd)."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2f, a2, b2, cfg):
        require_cuda(x, W1)
        ctx.set_materialize_grads(False)
        dev = x.device
        x = x.contiguous()
        B, N, H = x.shape
        bf = cfg.precision == 'bf16'
        d_mid, d_out = cfg.drops
        d = _lib.FfnBlock()
        d.precision = 1 if bf else 0
        d.M, d.H, d.F = B * N, H, W1.shape[0]
        d.residual = 1 if cfg.residual else 0
        d.eps = cfg.eps
        st = d_mid.state if d_mid.active else (d_out.state if d_out.active else None)
        if st is not None:
            d.rng_state = st.data_ptr()
            if d_mid.active:
                d.p_mid, d.salt_mid = d_mid.p, d_mid.salt
            if d_out.active:
                d.p_out, d.salt_out = d_out.p, d_out.salt
        x16 = None
        if bf:
            x16 = cfg.x16
            d.w16_1, d.w16_2 = cfg.w16['w1'].data_ptr(), cfg.w16['w2'].data_ptr()
        d.x, d.x16 = x.data_ptr(), _ptr(x16)
        d.W1, d.b1, d.W2, d.b2 = W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2f.data_ptr()
        d.ln_a, d.ln_b = _ptr(a2), _ptr(b2)
        fwd_bytes, bwd_bytes = _lib.workspace_bytes(d)
        ws = torch.empty(fwd_bytes, dtype=torch.uint8, device=dev)
        out = torch.empty((B, N, H), dtype=torch.float32, device=dev)
        out16 = torch.empty((B, N, H), dtype=torch.bfloat16, device=dev) if bf else None
        d.out, d.out16, d.workspace = out.data_ptr(), _ptr(out16), ws.data_ptr()
        d.stream = _lib.stream()
        _lib.call_block('mmnas_ffn_ln_fwd', d)
        ctx.desc, ctx.bwd_bytes, ctx.bf, ctx.shape = d, bwd_bytes, bf, (B, N, H)
        ctx.params = (W1, b1, W2, b2f, a2, b2)
        ctx.save_for_backward(x, ws, x16)
        ctx.w16 = cfg.w16
        if bf:
            ctx.mark_non_differentiable(out16)
            return out, out16
        return out, None

    @staticmethod
    def backward(ctx, dout, _unused=None):
        if dout is None:
            return (None,) * 8
        d = ctx.desc
        x, ws, x16 = ctx.saved_tensors
        W1, b1, W2, b2f, a2, b2 = ctx.params
        dev = x.device
        dout = dout.contiguous()
        dx = torch.empty(ctx.shape, dtype=torch.float32, device=dev)
        bws = torch.empty(ctx.bwd_bytes, dtype=torch.uint8, device=dev)
        own = (W1, b1, W2, b2f, a2, b2)
        sinks = _direct(own)
        direct = sinks is not None
        if not direct:
            f32 = dict(dtype=torch.float32, device=dev)
            sinks = [torch.empty_like(W1, dtype=torch.float32), torch.empty((d.F,), **f32),
                     torch.empty_like(W2, dtype=torch.float32), torch.empty((d.H,), **f32), None, None]
            if a2 is not None:
                gab = torch.empty((2, d.H), **f32)
                sinks[4], sinks[5] = gab[0], gab[1]
        d.accumulate_grads = 1 if direct else 0
        d.dW1, d.db1, d.dW2, d.db2 = (sinks[i].data_ptr() for i in range(4))
        d.dln_a, d.dln_b = _ptr(sinks[4]), _ptr(sinks[5])
        d.dout, d.dx, d.bwd_workspace = dout.data_ptr(), dx.data_ptr(), bws.data_ptr()
        d.stream = _lib.stream()
        d.side_stream = _side_stream(dev, ctx.bf)
        _lib.call_block('mmnas_ffn_ln_bwd', d)
        if direct:
            runtime.notify_grads(own)
            return (dx, None, None, None, None, None, None, None)
        return (dx, sinks[0], sinks[1], sinks[2], sinks[3], sinks[4], sinks[5], None)


# ----------------------------------------------------------------------------------------------------------
# stem: y_in = imgfeat_linear(frcn_feat), y_mask = make_mask(frcn_feat)   (full_vqa.py:91,101,113-114)
# ----------------------------------------------------------------------------------------------------------
class StemImageFn(Function):
    """One pass over the [B, N, 2048] region features produces their bf16 copy and the padding mask; the projection
    and its weight gradient run on the tcgen05 GEMM (bf16 arm) or the FFMA GEMM (fp32 arm)."""

    @staticmethod
    def forward(ctx, feat, W, b, precision):
        require_cuda(feat, W)
        ctx.set_materialize_grads(False)
        dev = feat.device
        feat = feat.contiguous()
        B, N, Fin = feat.shape
        H = W.shape[0]
        M = B * N
        bf = precision == 'bf16'
        mask = torch.empty((B, N), dtype=torch.uint8, device=dev)
        y = _empty((B, N, H), torch.float32, dev)
        feat16 = None
        if feat.dtype == torch.bfloat16:     # compact loader format: the features arrive as bf16 (what the bf16 arm
            feat16 = feat.view(M, Fin)       # casts them to anyway), only the mask is left to compute
            K.rowmask_bf16(feat16, mask, M, Fin)
            if not bf:
                feat = feat.float()
        if bf:
            if feat16 is None:
                feat16 = _empty((M, Fin), torch.bfloat16, dev)
                K.cast_rowmask(feat, feat16, mask, M, Fin)
            w16 = K.cast_bf16(W.detach())
            K.gemm_bf16(M, H, Fin, feat16, Fin, 0, w16, Fin, 0, y, H, bias=b)
        else:
            if feat16 is None:
                K.cast_rowmask(feat, None, mask, M, Fin)
            K.gemm_f32(M, H, Fin, feat, Fin, 1, W, 1, Fin, y, H, bias=b)
        ctx.bf, ctx.dims = bf, (M, Fin, H)
        ctx.params = (W, b)
        ctx.save_for_backward(feat16 if bf else feat)
        return y, mask.view(B, 1, 1, N).view(torch.bool)      # bool output: non-differentiable by type

    @staticmethod
    def backward(ctx, dy, _dmask=None):
        if dy is None:
            return None, None, None, None
        M, Fin, H = ctx.dims
        W, b = ctx.params
        x, = ctx.saved_tensors
        dev = dy.device
        dy = dy.contiguous().view(M, H)
        s_W, s_b = _Sink([W], dev), _Sink([b], dev)
        s_b.prepare(False)
        K.colsum(dy, M, H, H, s_b.buf, accumulate=s_b.direct)
        if ctx.bf:
            dy16 = K.cast_bf16(dy)
            sk = _split_k(H, Fin, M)
            s_W.prepare(zero=sk > 1)
            K.gemm_bf16(H, Fin, M, dy16, H, 1, x, Fin, 1, s_W.buf, Fin, split_k=sk, accumulate=s_W.direct and sk == 1)
        else:
            s_W.prepare(False)
            K.gemm_f32(H, Fin, M, dy, 1, H, x, Fin, 1, s_W.buf, Fin, accumulate=s_W.direct)
        return None, s_W.grads()[0], s_b.grads()[0], None


# ----------------------------------------------------------------------------------------------------------
# callers' dense layers: the Linear(+ReLU+dropout) layers of AttFlat and of the task heads
# (modules.py:13-41,59-85; full_vqa.py:105-109; full_vgd.py:105-112; full_itm.py:109-110) on the library GEMMs
# ----------------------------------------------------------------------------------------------------------
def _pad32(n):
    return (n + 31) // 32 * 32


class LinearFn(Function):
    """y = dropout(relu(x W^T + b)) (ReLU / dropout optional) over the last dimension of x.

    bf16 arm: tcgen05 GEMM with the bias / ReLU / dropout epilogue; backward = one masked cast, a column sum, a split-K
    weight-gradient GEMM and a dgrad GEMM.  fp32 arm: the FFMA GEMM.  The tensor-core GEMM wants an output width that
    is a multiple of 32: other widths (3129 answers, the 1- and 4-wide scoring layers) run on zero-padded weight rows
    and a padded output pitch, and hand back the leading columns."""

    @staticmethod
    def forward(ctx, x, W, b, relu, drop, precision):
        require_cuda(x, W)
        ctx.set_materialize_grads(False)
        dev = x.device
        N, Kin = W.shape
        shape = x.shape
        x2 = x.reshape(-1, Kin).contiguous()
        M = x2.shape[0]
        # the tensor-core GEMM needs every GEMM 'N' of forward / dgrad / wgrad to be a multiple of 32: the output width is
        # padded, an input width that is not (toy configurations only) takes the FFMA GEMM
        bf = precision == 'bf16' and Kin % 32 == 0
        Np = _pad32(N) if bf else N
        y = torch.empty((M, Np), dtype=torch.float32, device=dev)
        if bf:
            x16 = getattr(x, '_mmnas_bf16', None)
            x16 = x16.reshape(M, Kin) if (x16 is not None and x16.numel() == x2.numel()) else K.cast_bf16(x2)
            if Np == N:
                w16, bp = K.cast_bf16(W.detach()), b
            else:
                w16 = torch.zeros((Np, Kin), dtype=torch.bfloat16, device=dev)
                K.cast_bf16(W.detach(), w16[:N])
                bp = None
                if b is not None:
                    bp = torch.zeros(Np, dtype=torch.float32, device=dev)
                    bp[:N].copy_(b.detach())
            K.gemm_bf16(M, Np, Kin, x16, Kin, 0, w16, Kin, 0, y, Np, bias=bp, relu=relu, drop=drop)
            ctx.save_for_backward(x16, w16, y if relu else None)
        else:
            K.gemm_f32(M, N, Kin, x2, Kin, 1, W, 1, Kin, y, N, bias=b, epilogue=(2 if drop.active else 1) if relu else 0,
                       drop=drop if relu else K.NO_DROP)
            ctx.save_for_backward(x2, W, y if relu else None)
        if drop.active and not relu:
            raise NotImplementedError('dropout without ReLU is not used by any caller layer')
        ctx.meta = (bf, M, N, Np, Kin, shape, relu, (1.0 / (1.0 - drop.p)) if drop.active else 1.0)
        ctx.params = (W, b)
        out = y[:, :N] if Np != N else y
        return out.reshape(shape[:-1] + (N,))

    @staticmethod
    def backward(ctx, dy):
        if dy is None:
            return (None,) * 6
        bf, M, N, Np, Kin, shape, relu, scale = ctx.meta
        xs, ws, y = ctx.saved_tensors
        W, b = ctx.params
        dev = dy.device
        dy2 = dy.reshape(M, N)
        # gradient w.r.t. the pre-activation: the saved output is 0 exactly where ReLU or dropout cut the element
        if relu:
            dpre = dy2 * (y[:, :N] > 0)
            if scale != 1.0:
                dpre = dpre * scale
        else:
            dpre = dy2
        sinks = _direct((W, b)) if b is not None else _direct((W,))
        direct = sinks is not None
        gW = sinks[0] if direct else torch.empty_like(W, dtype=torch.float32)
        gb = None
        if b is not None:
            gb = sinks[1] if direct else torch.empty_like(b, dtype=torch.float32)
        dx = torch.empty((M, Kin), dtype=torch.float32, device=dev)
        if bf:
            d16 = torch.zeros((M, Np), dtype=torch.bfloat16, device=dev) if Np != N else torch.empty((M, N), dtype=torch.bfloat16, device=dev)
            d16[:, :N].copy_(dpre)
            if b is not None:
                K.colsum(d16, M, N, Np, gb, accumulate=direct)
            sk = _split_k(N, Kin, M)
            if sk > 1 and not direct:
                gW.zero_()
            K.gemm_bf16(N, Kin, M, d16, Np, 1, xs, Kin, 1, gW, Kin, split_k=sk, accumulate=direct and sk == 1)
            K.gemm_bf16(M, Kin, Np, d16, Np, 0, ws, Kin, 1, dx, Kin)
        else:
            dpre = dpre.contiguous()
            if b is not None:
                K.colsum(dpre, M, N, N, gb, accumulate=direct)
            K.gemm_f32(N, Kin, M, dpre, 1, N, xs, Kin, 1, gW, Kin, accumulate=direct)
            K.gemm_f32(M, Kin, N, dpre, N, 1, W, Kin, 1, dx, Kin)
        if direct:
            runtime.notify_grads((W, b) if b is not None else (W,))
            gW = gb = None
        return dx.reshape(shape), gW, gb, None, None, None


class LSTMFn(Function):
    """nn.LSTM(E -> H, one layer, batch_first, zero initial state), output sequence only (full_vqa.py:68-74,94-95),
    bf16 arm.  The input projection of all steps is one tensor-core GEMM; the recurrence is ONE kernel per direction
    (csrc/lstm.cu) instead of cuDNN's two launches per step; weight / input gradients are GEMMs over the saved gate
    gradients, the weight-gradient ones on the side stream next to the input-gradient one.  Buffers are sequence-major
    inside, the output batch-first as nn.LSTM returns it (with its bf16 copy attached for the first encoder block)."""

    @staticmethod
    def forward(ctx, emb, w_ih, w_hh, b_ih, b_hh):
        require_cuda(emb, w_ih, w_hh)
        ctx.set_materialize_grads(False)
        dev = emb.device
        B, T, E = emb.shape
        H = w_hh.shape[1]
        Ep = _pad32(E)
        TB = T * B
        x16 = torch.zeros((TB, Ep), dtype=torch.bfloat16, device=dev) if Ep != E else torch.empty((TB, E), dtype=torch.bfloat16, device=dev)
        x16.view(T, B, Ep)[:, :, :E].copy_(emb.detach().transpose(0, 1))
        wih16 = torch.zeros((4 * H, Ep), dtype=torch.bfloat16, device=dev) if Ep != E else torch.empty((4 * H, E), dtype=torch.bfloat16, device=dev)
        wih16[:, :E].copy_(w_ih.detach())
        whh16 = K.cast_bf16(w_hh.detach())
        bias = (b_ih.detach() + b_hh.detach()) if b_ih is not None else None
        xw = torch.empty((TB, 4 * H), dtype=torch.float32, device=dev)
        K.gemm_bf16(TB, 4 * H, Ep, x16, Ep, 0, wih16, Ep, 0, xw, 4 * H, bias=bias)
        ws = torch.empty(K.lstm_workspace(T, B, H), dtype=torch.uint8, device=dev)
        out = torch.empty((B, T, H), dtype=torch.float32, device=dev)
        out16 = torch.empty((B, T, H), dtype=torch.bfloat16, device=dev)
        K.lstm_fwd(T, B, H, xw, whh16, out, out16, ws)
        out._mmnas_bf16 = out16
        ctx.save_for_backward(x16, wih16, whh16, ws)
        ctx.meta = (T, B, E, Ep, H)
        ctx.params = (w_ih, w_hh, b_ih, b_hh)
        return out

    @staticmethod
    def backward(ctx, dout):
        if dout is None:
            return (None,) * 5
        T, B, E, Ep, H = ctx.meta
        x16, wih16, whh16, ws = ctx.saved_tensors
        w_ih, w_hh, b_ih, b_hh = ctx.params
        dev = dout.device
        TB = T * B
        K.lstm_bwd(T, B, H, dout.contiguous().float(), whh16, ws)
        al = lambda v: (v + 255) & ~255                      # workspace layout of csrc/lstm.cu
        o_dg = al((TB + B) * H * 2) + al(TB * 4 * H * 4) + al(TB * H * 4)
        h16 = ws[:TB * H * 2].view(torch.bfloat16).view(TB, H)                            # h_{t-1}, t = 0..T-1
        dg = ws[o_dg:o_dg + TB * 4 * H * 2].view(torch.bfloat16).view(TB, 4 * H)
        hh, bb = _Sink((w_hh,), dev), (_Sink((b_ih,), dev), _Sink((b_hh,), dev)) if b_ih is not None else ()
        ih = _Sink((w_ih,), dev) if Ep == E else None
        # destinations first, on the main stream: inside `with fork` only the library's launches move to the side stream
        sk_hh, sk_ih = _split_k(4 * H, H, TB), _split_k(4 * H, Ep, TB)
        hh.prepare(zero=sk_hh > 1)
        if ih is not None:
            ih.prepare(zero=sk_ih > 1)
            g_ih = ih.buf
        else:
            g_ih = (torch.zeros if sk_ih > 1 else torch.empty)((4 * H, Ep), dtype=torch.float32, device=dev)
        for sink in bb:
            sink.prepare(zero=False)
        fork = _Fork(dev, True)
        with fork:                                           # weight gradients next to the input-gradient GEMM
            K.gemm_bf16(4 * H, H, TB, dg, 4 * H, 1, h16, H, 1, hh.buf, H, split_k=sk_hh, accumulate=hh.direct and sk_hh == 1)
            K.gemm_bf16(4 * H, Ep, TB, dg, 4 * H, 1, x16, Ep, 1, g_ih, Ep, split_k=sk_ih,
                        accumulate=ih is not None and ih.direct and sk_ih == 1)
            for sink in bb:
                K.colsum(dg, TB, 4 * H, 4 * H, sink.buf, accumulate=sink.direct)
        dxp = torch.empty((TB, Ep), dtype=torch.float32, device=dev)
        K.gemm_bf16(TB, Ep, 4 * H, dg, 4 * H, 0, wih16, Ep, 1, dxp, Ep)
        demb = dxp.view(T, B, Ep)[:, :, :E].transpose(0, 1)
        fork.join()
        g_w_ih = ih.grads()[0] if ih is not None else g_ih[:, :E]
        g_b = tuple(sink.grads()[0] for sink in bb) if bb else (None, None)
        return demb, g_w_ih, hh.grads()[0], g_b[0], g_b[1]


def lstm(emb, mod):
    """`mod(emb)[0]` for a one-layer batch_first nn.LSTM: the cluster kernels of csrc/lstm.cu in the bf16 arm on CUDA
    (H = 256 / 512, at most 256 rows), torch's LSTM otherwise — the fp32 parity arm, other shapes, and devices that
    cannot schedule a 16-CTA cluster (the library reports MMNAS_ERR_UNSUPPORTED; said once, then torch's LSTM is used)."""
    H = mod.hidden_size
    if (emb.is_cuda and runtime.get_precision() == 'bf16' and mod.num_layers == 1 and mod.batch_first
            and not mod.bidirectional and mod.proj_size == 0 and H in (256, 512) and emb.shape[0] <= 256
            and runtime.native_lstm):
        try:
            return LSTMFn.apply(emb, mod.weight_ih_l0, mod.weight_hh_l0, mod.bias_ih_l0 if mod.bias else None,
                                mod.bias_hh_l0 if mod.bias else None)
        except _lib.MMnasLibraryError as e:
            if '(-2)' not in str(e):
                raise
            import warnings
            warnings.warn('mmnas_b200: LSTM kernels unsupported on this device (%s); using torch.nn.LSTM' % e)
            runtime.native_lstm = False
    return mod(emb)[0]


class AddLayerNormFn(Function):
    """LayerNorm(a + b) over the last dimension (full_vqa.py:107-108: proj_norm(attflat_x + attflat_y); full_vgd.py:108-109
    with a broadcast over the regions): the residual + LayerNorm kernel with `a` as the residual input."""

    @staticmethod
    def forward(ctx, a, b, a2, b2, eps):
        require_cuda(a, b)
        shape = b.shape
        H = shape[-1]
        z = b.reshape(-1, H).to(torch.float32).clone(memory_format=torch.contiguous_format)
        res = a.expand(shape).reshape(-1, H).to(torch.float32).contiguous()
        rows = z.shape[0]
        out = torch.empty_like(z)
        mean = torch.empty(rows, dtype=torch.float32, device=z.device)
        sigma = torch.empty(rows, dtype=torch.float32, device=z.device)
        K.ln_residual_fwd(rows, H, res, z, a2, b2, eps, out, None, mean, sigma)
        ctx.save_for_backward(z, mean, sigma, a2)
        ctx.eps, ctx.shape, ctx.a_shape, ctx.params = eps, shape, a.shape, (a2, b2)
        return out.view(shape)

    @staticmethod
    def backward(ctx, dout):
        z, mean, sigma, a2 = ctx.saved_tensors
        rows, H = z.shape
        dout = dout.reshape(rows, H).contiguous()
        dz = torch.empty_like(z)
        sinks = _direct(ctx.params)
        direct = sinks is not None
        da2, db2 = sinks if direct else (torch.zeros_like(a2), torch.zeros_like(a2))
        K.ln_residual_bwd(rows, H, dout, z, mean, sigma, a2, ctx.eps, dz, None, da2, db2)
        dz = dz.view(ctx.shape)
        da = dz
        if tuple(ctx.a_shape) != tuple(ctx.shape):          # `a` was broadcast over leading dimensions: sum them back
            da = dz.sum_to_size(ctx.a_shape)
        if direct:
            runtime.notify_grads(ctx.params)
            return da, dz, None, None, None
        return da, dz, da2, db2, None


# ----------------------------------------------------------------------------------------------------------
# stand-alone LayerNorm (modules.py:44-56) — same kernel with no residual / dropout
# ----------------------------------------------------------------------------------------------------------
class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, a2, b2, eps):
        require_cuda(x)
        shape = x.shape
        H = shape[-1]
        z = x.reshape(-1, H).to(torch.float32).clone(memory_format=torch.contiguous_format)
        rows = z.shape[0]
        out = torch.empty_like(z)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        sigma = torch.empty(rows, dtype=torch.float32, device=x.device)
        K.ln_residual_fwd(rows, H, None, z, a2, b2, eps, out, None, mean, sigma)
        ctx.save_for_backward(z, mean, sigma, a2)
        ctx.eps, ctx.shape = eps, shape
        return out.view(shape)

    @staticmethod
    def backward(ctx, dout):
        z, mean, sigma, a2 = ctx.saved_tensors
        rows, H = z.shape
        dout = dout.reshape(rows, H).contiguous()
        dz = torch.empty_like(z)
        da2 = torch.zeros_like(a2)
        db2 = torch.zeros_like(a2)
        K.ln_residual_bwd(rows, H, dout, z, mean, sigma, a2, ctx.eps, dz, None, da2, db2)
        return dz.view(ctx.shape), da2, db2, None


# ----------------------------------------------------------------------------------------------------------
# supernet mixed-op, 'full' / 'two' modes  (mixed.py:60-68): out = sum_k gate_k * o_k, inactive o_k detached
# ----------------------------------------------------------------------------------------------------------
class MixedSumFn(Function):
    """forward(gate, n_active, o_0 ... o_{K-1}) with the first n_active candidates taking gradient."""

    @staticmethod
    def forward(ctx, gate, n_active, *outs):
        require_cuda(gate, *outs)
        outs = [o.contiguous() for o in outs]
        out = torch.empty_like(outs[0])
        K.mixed_accum(outs, gate, out)
        ctx.n_active = n_active
        ctx.save_for_backward(gate, *outs)
        return out

    @staticmethod
    def backward(ctx, dout):
        gate, *outs = ctx.saved_tensors
        dout = dout.contiguous()
        gate_grad = torch.empty(len(outs), dtype=torch.float32, device=gate.device)
        d_outs = [torch.empty_like(o) if (i < ctx.n_active and ctx.needs_input_grad[2 + i]) else None
                  for i, o in enumerate(outs)]
        K.mixed_alpha_dot(outs, gate, dout, gate_grad, d_outs)
        return (gate_grad, None) + tuple(d_outs)
