"""Thin Python wrappers over the C ABI (one function per entry point).  They only translate torch tensors
into raw pointers / pitches; no arithmetic happens here."""
import ctypes

import torch

from . import _lib
from ._lib import call, ptr, stream

F32, BF16 = 0, 1


class Drop:
    """One dropout site: device rng state {seed, step}, a salt unique to the site+call, probability p."""
    __slots__ = ('state', 'salt', 'p')

    def __init__(self, state=None, salt=0, p=0.0):
        self.state, self.salt, self.p = state, int(salt) & 0xFFFFFFFFFFFFFFFF, float(p)

    @property
    def active(self):
        return self.state is not None and self.p > 0.0

    def args(self):
        if not self.active:
            return None, 0, 0.0
        return self.state.data_ptr(), self.salt, self.p


NO_DROP = Drop()


def _code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError('unsupported dtype %s' % t.dtype)


def gemm_f32(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias=None, epilogue=0, accumulate=False, aux=None,
             ld_aux=0, aux_scale=1.0, drop=NO_DROP):
    st, salt, p = drop.args()
    call('mmnas_gemm_f32', M, N, K, ptr(A), a_rs, a_cs, ptr(B), b_rs, b_cs, ptr(C), ldc, ptr(bias), epilogue,
         int(accumulate), ptr(aux), ld_aux, aux_scale, st, salt, p, stream())


def gemm_bf16(M, N, K, A, lda, a_mn, B, ldb, b_mn, C, ldc, bias=None, relu=False, accumulate=False, aux=None,
              ld_aux=0, aux_scale=1.0, split_k=1, drop=NO_DROP):
    st, salt, p = drop.args()
    call('mmnas_gemm_bf16', M, N, K, ptr(A), lda, int(a_mn), ptr(B), ldb, int(b_mn), ptr(C), ldc,
         int(C.dtype == torch.bfloat16), ptr(bias), int(relu), int(accumulate), ptr(aux), ld_aux, aux_scale,
         split_k, st, salt, p, stream())


def attn_fwd(B, heads, Nq, Nk, q, ldq, k, ldk, v, ldv, kmask, bias, o, ldo, scale, drop=NO_DROP, head_dim=64):
    st, salt, p = drop.args()
    call('mmnas_attn_fwd', _code(o), B, heads, Nq, Nk, head_dim, q, ldq, k, ldk, v, ldv, ptr(kmask), ptr(bias),
         ptr(o), ldo, scale, st, salt, p, stream())


def attn_bwd(B, heads, Nq, Nk, q, ldq, k, ldk, v, ldv, kmask, bias, o, ldo, dout, lddo, dq, lddq, dk, lddk, dv, lddv,
             dbias, scale, drop=NO_DROP, head_dim=64):
    st, salt, p = drop.args()
    call('mmnas_attn_bwd', _code(o), B, heads, Nq, Nk, head_dim, q, ldq, k, ldk, v, ldv, ptr(kmask), ptr(bias),
         ptr(o), ldo, ptr(dout), lddo, dq, lddq, dk, lddk, dv, lddv, ptr(dbias), scale, st, salt, p, stream())


def relbias_fwd(B, N, heads, R, rel, g4, Wy, by, Wr, br, bias, mode=0):
    """mode 0 = fp32 arithmetic, 1 = tensor-core arithmetic (bf16 arm)"""
    call('mmnas_relbias_fwd', mode, B, N, heads, R, ptr(rel), ptr(g4), ptr(Wy), ptr(by), ptr(Wr), ptr(br), ptr(bias), stream())


def relbias_bwd(B, N, heads, R, rel, g4, Wy, by, Wr, br, dbias, drel, dWy, dby, dWr, dbr, mode=0):
    call('mmnas_relbias_bwd', mode, B, N, heads, R, ptr(rel), ptr(g4), ptr(Wy), ptr(by), ptr(Wr), ptr(br), ptr(dbias),
         ptr(drel), ptr(dWy), ptr(dby), ptr(dWr), ptr(dbr), stream())


def ln_residual_fwd(rows, H, x, branch, gamma, beta, eps, out, out16, mean, sigma, drop=NO_DROP):
    st, salt, p = drop.args()
    call('mmnas_ln_residual_fwd', rows, H, ptr(x), ptr(branch), ptr(gamma), ptr(beta), eps, ptr(out), ptr(out16),
         ptr(mean), ptr(sigma), st, salt, p, stream())


def ln_residual_bwd(rows, H, dout, z, mean, sigma, gamma, eps, dz, dbranch, dgamma, dbeta, drop=NO_DROP):
    st, salt, p = drop.args()
    dt = F32 if dbranch is None else _code(dbranch)
    call('mmnas_ln_residual_bwd', rows, H, ptr(dout), ptr(z), ptr(mean), ptr(sigma), ptr(gamma), eps, ptr(dz),
         ptr(dbranch), dt, ptr(dgamma), ptr(dbeta), st, salt, p, stream())


def mixed_accum(outs, gate, out):
    call('mmnas_mixed_accum', len(outs), _lib.ptr_array(outs), ptr(gate), ptr(out), out.numel(), stream())


def mixed_alpha_dot(outs, gate, dout, gate_grad, d_outs):
    call('mmnas_mixed_alpha_dot', len(outs), _lib.ptr_array(outs), ptr(gate), ptr(dout), ptr(gate_grad),
         _lib.ptr_array(d_outs), dout.numel(), stream())


def cast_bf16(src, dst=None):
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    call('mmnas_cast_f32_to_bf16', ptr(src), ptr(dst), src.numel(), stream())
    return dst


def cast_multi(table, n_chunks):
    call('mmnas_cast_multi', ptr(table), n_chunks, stream())


def colsum(x, rows, cols, ld, out, accumulate=False):
    call('mmnas_colsum', _code(x), ptr(x), rows, cols, ld, ptr(out), int(accumulate), stream())


def cast_rowmask(x, x16, mask, rows, cols):
    call('mmnas_cast_rowmask', ptr(x), ptr(x16), ptr(mask), rows, cols, stream())


SUMSQ_SCRATCH = 1280      # MMNAS_SUMSQ_SCRATCH floats


def sumsq(flat, out, scratch):
    assert scratch.numel() >= SUMSQ_SCRATCH and scratch.dtype == torch.float32
    call('mmnas_sumsq_f32', ptr(flat), flat.numel(), ptr(out), ptr(scratch), stream())


def clip_adam(table, n_chunks, sumsq_buf, lr, step_state, beta1, beta2, eps, max_norm):
    call('mmnas_clip_adam', ptr(table), n_chunks, ptr(sumsq_buf), ptr(lr), ptr(step_state), beta1, beta2, eps, max_norm,
         stream())


def rng_advance(state):
    call('mmnas_rng_advance', ptr(state), stream())


def box_geometry(boxes, pad_mask):
    """boxes [B,N,4] fp32, pad_mask uint8 [B,N] (1 = padded) or None -> g4 [B,N,N,4] (load_data_vqa.py:7-33,236-239)."""
    B, N = boxes.shape[0], boxes.shape[1]
    g4 = torch.empty((B, N, N, 4), dtype=torch.float32, device=boxes.device)
    call('mmnas_box_geometry', ptr(boxes), ptr(pad_mask), ptr(g4), B, N, stream())
    return g4


def rowmask_bf16(x16, mask, rows, cols):
    call('mmnas_rowmask_bf16', ptr(x16), ptr(mask), rows, cols, stream())


def lstm_workspace(T, B, H):
    n = ctypes.c_ulonglong(0)
    call('mmnas_lstm_workspace', T, B, H, ctypes.addressof(n))
    return int(n.value)


def lstm_fwd(T, B, H, xw, whh16, out, out16, workspace):
    call('mmnas_lstm_fwd', T, B, H, ptr(xw), ptr(whh16), ptr(out), ptr(out16), ptr(workspace), stream())


def lstm_bwd(T, B, H, dout, whh16, workspace):
    call('mmnas_lstm_bwd', T, B, H, ptr(dout), ptr(whh16), ptr(workspace), stream())


def gemm_ln_bf16(M, N, K, A, lda, W, ldb, bias, x, gamma, beta, eps, z, out, out16, mean, sigma, drop=NO_DROP):
    """Fused projection + residual + LayerNorm (bf16 arm).  Raises MMnasLibraryError(-2) for unsupported shapes."""
    st, salt, p = drop.args()
    call('mmnas_gemm_ln_bf16', M, N, K, ptr(A), lda, ptr(W), ldb, ptr(bias), ptr(x), ptr(gamma), ptr(beta), eps, ptr(z), ptr(out),
         ptr(out16), ptr(mean), ptr(sigma), st, salt, p, stream())
