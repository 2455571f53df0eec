"""Kernel-level parity tests, called through the C ABI (ctypes) on the B200.  Each kernel is compared with
the same arithmetic written in float64 torch on identical seeded inputs (the per-kernel restatement of the
reference lines cited in include/mmnas_b200.h); tolerances are normwise."""
import math

import pytest
import torch

from tests.util import normwise

pytestmark = pytest.mark.gpu

DEV = 'cuda'


def K():
    from mmnas_b200 import kernels
    return kernels


def rnd(*shape, seed=0, dtype=torch.float32, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


# ------------------------------------------------------------------------------------------ GEMM fp32
@pytest.mark.parametrize('ta,tb', [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize('M,N,K_', [(200, 96, 130), (64, 64, 64), (37, 5, 11)])
def test_gemm_f32_layouts(ta, tb, M, N, K_):
    k = K()
    A = rnd(K_, M, seed=1) if ta else rnd(M, K_, seed=1)
    B = rnd(N, K_, seed=2) if tb else rnd(K_, N, seed=2)
    bias = rnd(N, seed=3)
    C = torch.empty(M, N, device=DEV)
    a_rs, a_cs = (1, M) if ta else (K_, 1)
    b_rs, b_cs = (1, K_) if tb else (N, 1)
    k.gemm_f32(M, N, K_, A, a_rs, a_cs, B, b_rs, b_cs, C, N, bias=bias, epilogue=1)
    Ad = (A.t() if ta else A).double()
    Bd = (B.t() if tb else B).double()
    ref = torch.relu(Ad @ Bd + bias.double())
    assert normwise(C, ref) < 2e-6
    # accumulate + aux-mask epilogue
    aux = rnd(M, N, seed=4)
    C2 = C.clone()
    k.gemm_f32(M, N, K_, A, a_rs, a_cs, B, b_rs, b_cs, C2, N, epilogue=3, accumulate=True, aux=aux, ld_aux=N, aux_scale=1.25)
    ref2 = ref + torch.where(aux.double() > 0, (Ad @ Bd) * 1.25, torch.zeros_like(ref))
    assert normwise(C2, ref2) < 2e-6


# ------------------------------------------------------------------------------------------ GEMM bf16 (tcgen05)
@pytest.mark.parametrize('a_mn,b_mn', [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize('M,N,K_', [(128, 128, 64), (200, 256, 192), (896, 512, 512), (300, 96, 1000)])
def test_gemm_bf16_layouts(a_mn, b_mn, M, N, K_):
    k = K()
    Kp = (K_ + 7) // 8 * 8
    Mp = (M + 7) // 8 * 8
    A = rnd(Kp if a_mn else M, Mp if a_mn else Kp, seed=1, dtype=torch.bfloat16)
    B = rnd(Kp if b_mn else N, N if b_mn else Kp, seed=2, dtype=torch.bfloat16)
    Ad = (A[:K_, :M].t() if a_mn else A[:, :K_]).double()
    Bd = (B[:K_] if b_mn else B[:, :K_].t()).double()
    ref = Ad @ Bd
    C = torch.full((M, N), float('nan'), device=DEV)
    k.gemm_bf16(M, N, K_, A, A.stride(0), a_mn, B, B.stride(0), b_mn, C, N)
    assert normwise(C, ref) < 1e-5
    # split-K accumulates into a zeroed fp32 buffer
    C3 = torch.zeros(M, N, device=DEV)
    k.gemm_bf16(M, N, K_, A, A.stride(0), a_mn, B, B.stride(0), b_mn, C3, N, split_k=3)
    assert normwise(C3, ref) < 1e-5


@pytest.mark.parametrize('a_mn,b_mn', [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize('M,N,K_', [(256, 256, 64), (1000, 512, 512), (6400, 256, 192), (300, 768, 1000)])
def test_gemm_bf16_cta_pair_path(a_mn, b_mn, M, N, K_, monkeypatch):
    """The cta_group::2 kernel (256x256 tiles over two CTAs) on every operand layout, ragged M (second CTA of the last
    pair partly / fully out of range), several tiles per cluster (accumulator double buffering), split-K."""
    monkeypatch.setenv('MMNAS_GEMM_PAIR', '1')
    k = K()
    Kp = (K_ + 7) // 8 * 8
    Mp = (M + 7) // 8 * 8
    A = rnd(Kp if a_mn else M, Mp if a_mn else Kp, seed=1, dtype=torch.bfloat16)
    B = rnd(Kp if b_mn else N, N if b_mn else Kp, seed=2, dtype=torch.bfloat16)
    Ad = (A[:K_, :M].t() if a_mn else A[:, :K_]).double()
    Bd = (B[:K_] if b_mn else B[:, :K_].t()).double()
    ref = Ad @ Bd
    C = torch.full((M, N), float('nan'), device=DEV)
    k.gemm_bf16(M, N, K_, A, A.stride(0), a_mn, B, B.stride(0), b_mn, C, N)
    assert normwise(C, ref) < 1e-5
    monkeypatch.setenv('MMNAS_GEMM_PAIR', '0')
    C1 = torch.empty(M, N, device=DEV)
    k.gemm_bf16(M, N, K_, A, A.stride(0), a_mn, B, B.stride(0), b_mn, C1, N)
    assert torch.equal(C, C1)                       # same products, same K order: the two tilings agree bit for bit
    monkeypatch.setenv('MMNAS_GEMM_PAIR', '1')
    C3 = torch.zeros(M, N, device=DEV)
    k.gemm_bf16(M, N, K_, A, A.stride(0), a_mn, B, B.stride(0), b_mn, C3, N, split_k=3)
    assert normwise(C3, ref) < 1e-5


def test_gemm_bf16_cta_pair_epilogues(monkeypatch):
    monkeypatch.setenv('MMNAS_GEMM_PAIR', '1')
    k = K()
    M, N, K_ = 2000, 512, 512
    A = rnd(M, K_, seed=1, dtype=torch.bfloat16)
    B = rnd(N, K_, seed=2, dtype=torch.bfloat16)
    bias = rnd(N, seed=3)
    ref = A.double() @ B.double().t()
    out16 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    k.gemm_bf16(M, N, K_, A, K_, 0, B, K_, 0, out16, N, bias=bias, relu=True)
    assert normwise(out16, torch.relu(ref + bias.double())) < 6e-3
    aux = rnd(M, N, seed=4, dtype=torch.bfloat16)
    k.gemm_bf16(M, N, K_, A, K_, 0, B, K_, 0, out16, N, aux=aux, ld_aux=N, aux_scale=0.5)
    assert normwise(out16, torch.where(aux.double() > 0, ref * 0.5, torch.zeros_like(ref))) < 6e-3
    C = rnd(M, N, seed=5)
    ref_acc = C.double() + ref
    k.gemm_bf16(M, N, K_, A, K_, 0, B, K_, 0, C, N, accumulate=True)
    assert normwise(C, ref_acc) < 1e-5
    st = torch.tensor([1234, 7], dtype=torch.int64, device=DEV)
    d = k.Drop(st, salt=99, p=0.1)
    o1, o2 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    k.gemm_bf16(M, N, K_, A, K_, 0, B, K_, 0, o1, N, drop=d)
    monkeypatch.setenv('MMNAS_GEMM_PAIR', '0')
    k.gemm_bf16(M, N, K_, A, K_, 0, B, K_, 0, o2, N, drop=d)
    assert torch.equal(o1, o2)                      # same dropout mask from either tiling


def test_gemm_bf16_epilogues():
    k = K()
    M, N, K_ = 333, 256, 512
    A = rnd(M, K_, seed=1, dtype=torch.bfloat16)
    B = rnd(N, K_, seed=2, dtype=torch.bfloat16)
    bias = rnd(N, seed=3)
    ref = A.double() @ B.double().t()
    out16 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    k.gemm_bf16(M, N, K_, A, K_, 0, B, K_, 0, out16, N, bias=bias, relu=True)
    assert normwise(out16, torch.relu(ref + bias.double())) < 6e-3
    aux = rnd(M, N, seed=4, dtype=torch.bfloat16)
    k.gemm_bf16(M, N, K_, A, K_, 0, B, K_, 0, out16, N, aux=aux, ld_aux=N, aux_scale=0.5)
    assert normwise(out16, torch.where(aux.double() > 0, ref * 0.5, torch.zeros_like(ref))) < 6e-3
    C = rnd(M, N, seed=5)
    ref_acc = C.double() + ref
    k.gemm_bf16(M, N, K_, A, K_, 0, B, K_, 0, C, N, accumulate=True)
    assert normwise(C, ref_acc) < 1e-5
    # column-slice output (pitch > N), as used for the fused QKV buffer
    big = torch.zeros(M, 3 * N, device=DEV, dtype=torch.bfloat16)
    k.gemm_bf16(M, N, K_, A, K_, 0, B, K_, 0, big[:, N:2 * N], 3 * N)
    assert normwise(big[:, N:2 * N], ref) < 6e-3
    assert float(big[:, :N].abs().max()) == 0 and float(big[:, 2 * N:].abs().max()) == 0


def test_gemm_bf16_dropout_matches_fp32_arm_mask():
    """Both arms hash (row*N+col) with the same key, so they drop the same elements."""
    k = K()
    M, N, K_ = 256, 128, 64
    A = torch.ones(M, K_, device=DEV)
    B = torch.ones(N, K_, device=DEV)
    st = torch.tensor([1234, 7], dtype=torch.int64, device=DEV)
    d = k.Drop(st, salt=99, p=0.1)
    C32 = torch.empty(M, N, device=DEV)
    k.gemm_f32(M, N, K_, A, K_, 1, B, 1, K_, C32, N, epilogue=2, drop=d)
    C16 = torch.empty(M, N, device=DEV)
    k.gemm_bf16(M, N, K_, A.bfloat16(), K_, 0, B.bfloat16(), K_, 0, C16, N, relu=True, drop=d)
    assert torch.equal(C32 == 0, C16 == 0)
    frac = float((C32 == 0).float().mean())
    assert abs(frac - 0.1) < 0.01
    kept = C32[C32 != 0]
    assert torch.allclose(kept, torch.full_like(kept, K_ / 0.9), rtol=1e-6)


# ------------------------------------------------------------------------------------------ LayerNorm tail
@pytest.mark.parametrize('rows,H', [(7, 64), (896, 512), (301, 256), (5, 1024)])
@pytest.mark.parametrize('residual', [True, False])
def test_ln_residual_fwd_bwd(rows, H, residual):
    k = K()
    x, br = rnd(rows, H, seed=1), rnd(rows, H, seed=2)
    a2, b2 = 1 + 0.1 * rnd(H, seed=3), 0.1 * rnd(H, seed=4)
    go = rnd(rows, H, seed=5)
    xd, bd, ad, bbd = (t.double().requires_grad_(True) for t in (x, br, a2, b2))
    z = xd + bd if residual else bd
    mu = z.mean(-1, keepdim=True)
    ref = ad * (z - mu) / (z.std(-1, keepdim=True) + 1e-6) + bbd
    ref.backward(go.double())
    out = torch.empty(rows, H, device=DEV)
    out16 = torch.empty(rows, H, device=DEV, dtype=torch.bfloat16)
    mean, sigma = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    zbuf = br.clone()
    k.ln_residual_fwd(rows, H, x if residual else None, zbuf, a2, b2, 1e-6, out, out16, mean, sigma)
    assert normwise(out, ref) < 2e-6
    assert normwise(out16, ref) < 5e-3
    assert normwise(zbuf, z) < 1e-6
    dz = torch.empty(rows, H, device=DEV)
    da, db = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    k.ln_residual_bwd(rows, H, go, zbuf, mean, sigma, a2, 1e-6, dz, None, da, db)
    assert normwise(dz, bd.grad) < 5e-6
    assert normwise(da, ad.grad) < 5e-6
    assert normwise(db, bbd.grad) < 5e-6


def test_ln_norm_off_and_dropout_consistency():
    k = K()
    rows, H = 64, 256
    x, br, go = rnd(rows, H, seed=1), torch.ones(rows, H, device=DEV), rnd(rows, H, seed=2)
    st = torch.tensor([5, 0], dtype=torch.int64, device=DEV)
    d = k.Drop(st, salt=3, p=0.25)
    out = torch.empty(rows, H, device=DEV)
    zbuf = br.clone()
    k.ln_residual_fwd(rows, H, x, zbuf, None, None, 1e-6, out, None, None, None, drop=d)
    mult = out - x                       # = dropout multiplier (branch was all ones)
    assert all(min(abs(u), abs(u - 1 / 0.75)) < 2e-3 for u in torch.unique(mult.round(decimals=3)).tolist())
    assert abs(float((mult == 0).float().mean()) - 0.25) < 0.02
    dz = torch.empty(rows, H, device=DEV)
    dbr = torch.empty(rows, H, device=DEV)
    k.ln_residual_bwd(rows, H, go, None, None, None, None, 1e-6, dz, dbr, None, None, drop=d)
    assert torch.equal(dz, go)
    assert torch.allclose(dbr, go * mult, rtol=1e-5, atol=1e-6)   # backward regenerates the same mask


# ------------------------------------------------------------------------------------------ attention core
def attn_ref(q, k, v, mask, bias, scale):
    B, Nq, I = q.shape
    h = I // 64
    qh, kh, vh = (t.view(B, -1, h, 64).transpose(1, 2) for t in (q, k, v))
    s = qh @ kh.transpose(-2, -1) * scale
    if bias is not None:
        s = bias + s
    if mask is not None:
        s = s.masked_fill(mask.view(B, 1, 1, -1).bool(), -1e9)
    p = torch.softmax(s, -1)
    return (p @ vh).transpose(1, 2).reshape(B, Nq, I)


@pytest.mark.parametrize('B,h,Nq,Nk,with_bias', [(2, 2, 7, 7, True), (3, 8, 100, 100, True), (3, 8, 100, 14, False),
                                                  (2, 4, 36, 50, False), (1, 1, 100, 128, False)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_attention_fwd_bwd(B, h, Nq, Nk, with_bias, dtype):
    k = K()
    I = 64 * h
    q, kk, v = rnd(B, Nq, I, seed=1, dtype=dtype), rnd(B, Nk, I, seed=2, dtype=dtype), rnd(B, Nk, I, seed=3, dtype=dtype)
    go = rnd(B, Nq, I, seed=4, dtype=dtype)
    mask = torch.zeros(B, Nk, dtype=torch.uint8, device=DEV)
    mask[0, Nk // 2:] = 1
    mask[B - 1, :] = 1                      # fully padded sample -> uniform attention
    bias = rnd(B, h, Nq, Nk, seed=5) if with_bias else None
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, kk, v))
    bd = bias.double().requires_grad_(True) if with_bias else None
    ref = attn_ref(qd, kd, vd, mask, bd, 0.125)
    ref.backward(go.double())
    o = torch.empty(B, Nq, I, device=DEV, dtype=dtype)
    k.attn_fwd(B, h, Nq, Nk, q.data_ptr(), I, kk.data_ptr(), I, v.data_ptr(), I, mask, bias, o, I, 0.125)
    tol = 2e-6 if dtype == torch.float32 else 8e-3
    assert normwise(o, ref) < tol
    if B > 1:
        assert normwise(o[B - 1].float(), v[B - 1].float().mean(0, keepdim=True).expand(Nq, I)) < (1e-5 if dtype == torch.float32 else 1e-2)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(kk), torch.empty_like(v)
    dbias = torch.empty(B, h, Nq, Nk, device=DEV) if with_bias else None
    k.attn_bwd(B, h, Nq, Nk, q.data_ptr(), I, kk.data_ptr(), I, v.data_ptr(), I, mask, bias, o, I, go, I,
               dq.data_ptr(), I, dk.data_ptr(), I, dv.data_ptr(), I, dbias, 0.125)
    tolb = 1e-5 if dtype == torch.float32 else 2e-2
    assert normwise(dq, qd.grad) < tolb
    assert normwise(dk, kd.grad) < tolb
    assert normwise(dv, vd.grad) < tolb
    if with_bias:
        assert normwise(dbias, bd.grad) < tolb


def test_attention_dropout_statistics_and_grad_consistency():
    k = K()
    B, h, N = 4, 2, 64
    I = 64 * h
    q, kk = torch.zeros(B, N, I, device=DEV), torch.zeros(B, N, I, device=DEV)   # uniform attention 1/N
    v = torch.ones(B, N, I, device=DEV)
    st = torch.tensor([42, 3], dtype=torch.int64, device=DEV)
    d = k.Drop(st, salt=11, p=0.1)
    o = torch.empty(B, N, I, device=DEV)
    k.attn_fwd(B, h, N, N, q.data_ptr(), I, kk.data_ptr(), I, v.data_ptr(), I, None, None, o, I, 0.125, drop=d)
    # each output = (#kept / N) / 0.9 ; mean over many rows ~ 1
    assert abs(float(o.mean()) - 1.0) < 0.01
    assert float(o.std()) > 0.0
    o2 = torch.empty_like(o)
    k.attn_fwd(B, h, N, N, q.data_ptr(), I, kk.data_ptr(), I, v.data_ptr(), I, None, None, o2, I, 0.125, drop=d)
    assert torch.equal(o, o2)             # same (state, salt) -> same mask
    st2 = torch.tensor([42, 4], dtype=torch.int64, device=DEV)
    k.attn_fwd(B, h, N, N, q.data_ptr(), I, kk.data_ptr(), I, v.data_ptr(), I, None, None, o2, I, 0.125,
               drop=k.Drop(st2, salt=11, p=0.1))
    assert not torch.equal(o, o2)         # next step -> fresh mask


# ------------------------------------------------------------------------------------------ RSA geometry bias
@pytest.mark.parametrize('mode', ['dense', 'geometry'])
@pytest.mark.parametrize('B,N,h', [(2, 7, 2), (3, 100, 8), (1, 37, 4)])
def test_relbias_fwd_bwd(mode, B, N, h):
    k = K()
    R = 64
    g4 = rnd(B, N, N, 4, seed=1)
    g4[0, N // 2:] = 0                    # zero-padded pairs, as the loader produces
    Wy, by = 0.5 * rnd(R, 4, seed=2), 0.1 * rnd(R, seed=3)
    Wr, br = 0.2 * rnd(h, R, seed=4), 0.1 * rnd(h, seed=5)
    if N > 8:
        # d log(r)/dr = 1/r makes the gradient sums ill-conditioned when some r sit just above the 1e-6 clamp
        # (fp32 evaluation of r then carries a relative error ~1e-7/r).  The small case above keeps the generic
        # regime (all branches: relu off, clamp, pass-through); the large cases use benign parameters: head 0
        # always inactive (relu off), the others well inside the pass-through region.
        Wr, br = 0.03 * rnd(h, R, seed=4), torch.ones(h, device=DEV)
        br[0] = -2.0
    go = rnd(B, h, N, N, seed=6)
    g4d, Wyd, byd, Wrd, brd = (t.double().requires_grad_(True) for t in (g4, Wy, by, Wr, br))
    e = torch.relu(g4d @ Wyd.t() + byd)
    if mode == 'dense':
        e = e.detach().requires_grad_(True)
    r = torch.relu(e @ Wrd.t() + brd).permute(0, 3, 1, 2)
    ref = torch.log(torch.clamp(r, min=1e-6))
    ref.backward(go.double())
    bias = torch.empty(B, h, N, N, device=DEV)
    dWr, dbr = torch.zeros_like(Wr), torch.zeros_like(br)
    if mode == 'dense':
        rel = e.detach().float().contiguous()
        k.relbias_fwd(B, N, h, R, rel, None, None, None, Wr, br, bias)
        drel = torch.empty_like(rel)
        k.relbias_bwd(B, N, h, R, rel, None, None, None, Wr, br, go, drel, None, None, dWr, dbr)
        assert normwise(drel, e.grad) < 1e-5
    else:
        k.relbias_fwd(B, N, h, R, None, g4, Wy, by, Wr, br, bias)
        dWy, dby = torch.zeros_like(Wy), torch.zeros_like(by)
        k.relbias_bwd(B, N, h, R, None, g4, Wy, by, Wr, br, go, None, dWy, dby, dWr, dbr)
        assert normwise(dWy, Wyd.grad) < 2e-5
        assert normwise(dby, byd.grad) < 2e-5
    # entries sitting exactly on the relu / clamp kink can flip between fp32 and fp64: compare away from it
    safe = (r.detach() > 1e-4) | (r.detach() == 0)
    assert normwise(bias[safe], ref[safe]) < 1e-5
    assert normwise(dWr, Wrd.grad) < 2e-5
    assert normwise(dbr, brd.grad) < 2e-5


# ------------------------------------------------------------------------------------------ mixed-op, helpers
@pytest.mark.parametrize('Kc', [2, 4])
def test_mixed_accum_and_alpha_dot(Kc):
    k = K()
    n = 64 * 100 * 256
    outs = [rnd(n, seed=10 + i) for i in range(Kc)]
    gate = torch.zeros(Kc, device=DEV)
    gate[1] = 1.0
    out = torch.empty(n, device=DEV)
    k.mixed_accum(outs, gate, out)
    assert torch.equal(out, outs[1])          # one-hot gate: exact
    gate2 = rnd(Kc, seed=3)
    k.mixed_accum(outs, gate2, out)
    ref = sum(g.double() * o.double() for g, o in zip(gate2, outs))
    assert normwise(out, ref) < 1e-6
    dout = rnd(n, seed=4)
    gg = torch.empty(Kc, device=DEV)
    d_outs = [torch.empty(n, device=DEV), None] + [None] * (Kc - 2)
    k.mixed_alpha_dot(outs, gate2, dout, gg, d_outs)
    ref_gg = torch.stack([(o.double() * dout.double()).sum() for o in outs])
    assert normwise(gg, ref_gg) < 1e-5
    assert normwise(d_outs[0], gate2[0].double() * dout.double()) < 1e-6


def test_cast_and_colsum():
    k = K()
    x = rnd(1000, 513, seed=1)
    x16 = k.cast_bf16(x)
    assert torch.equal(x16, x.bfloat16())
    out = torch.empty(513, device=DEV)
    k.colsum(x, 1000, 513, 513, out)
    assert normwise(out, x.double().sum(0)) < 1e-5
    k.colsum(x16, 1000, 513, 513, out)
    assert normwise(out, x16.double().sum(0)) < 1e-5


def test_bad_arguments_raise():
    from mmnas_b200._lib import MMnasLibraryError
    k = K()
    q = rnd(1, 8, 32)
    o = torch.empty_like(q)
    with pytest.raises(MMnasLibraryError):     # head dim 32 is not implemented: loud, no fallback
        k.attn_fwd(1, 1, 8, 8, q.data_ptr(), 32, q.data_ptr(), 32, q.data_ptr(), 32, None, None, o, 32, 0.1, head_dim=32)
    with pytest.raises(MMnasLibraryError):
        k.gemm_bf16(8, 30, 8, q.bfloat16(), 8, 0, q.bfloat16(), 8, 0, o, 30)   # N % 32 != 0


def test_fused_clip_adam_matches_torch():
    """mmnas_sumsq_f32 + mmnas_clip_adam against clip_grad_norm_(1.0) + torch.optim.Adam(betas=(.9,.98), eps=1e-9) — the
    step tail of train_vqa.py:309-311 — over 5 steps, incl. an odd-sized tensor (tail path) and a tiny one."""
    from mmnas_b200.engine import FlatGrads, WarmupAdam
    torch.manual_seed(0)
    shapes = [(512, 512), (3129,), (2048, 512), (7,), (64, 4)]
    ps_a = [torch.nn.Parameter(torch.randn(*s, device=DEV)) for s in shapes]
    ps_b = [torch.nn.Parameter(p.detach().clone()) for p in ps_a]
    fg = FlatGrads(ps_a)
    fused = WarmupAdam(ps_a, lr_base=1e-2, epoch_steps=2, flat_grads=fg)
    assert fused.fused is not None
    ref = torch.optim.Adam(ps_b, lr=0.0, betas=(0.9, 0.98), eps=1e-9)
    for step in range(5):
        fg.zero()
        gs = [torch.randn(*s, device=DEV) * (3.0 if step % 2 else 0.01) for s in shapes]   # clipped / not clipped
        for p, q, g in zip(ps_a, ps_b, gs):
            p.grad.copy_(g)
            q.grad = g.clone()
        fused.set_lr()
        fused.clip_and_step()
        torch.nn.utils.clip_grad_norm_(ps_b, 1.0)
        for grp in ref.param_groups:
            grp['lr'] = fused._rate
        ref.step()
    for p, q in zip(ps_a, ps_b):
        assert normwise(p, q) < 2e-6
    assert abs(fused._rate - 0.75e-2) < 1e-12 and fused.rate(1) == 2.5e-3      # step 5 of 2-step 'epochs': 3/4 warm-up


def test_sumsq_is_bit_reproducible_and_exact_enough():
    """The gradient-norm reduction feeds the clip coefficient of every replica: it must not depend on block
    scheduling (no float atomics), and it must match a float64 sum."""
    k = K()
    torch.manual_seed(1)
    x = torch.randn(58_000_004 // 4 * 4, device=DEV) * 0.01          # about the size of the MMnas-VQA gradient buffer
    out = torch.zeros(1, device=DEV)
    scratch = torch.zeros(k.SUMSQ_SCRATCH, device=DEV)
    vals = set()
    for _ in range(20):
        k.sumsq(x, out, scratch)
        vals.add(float(out))
    assert len(vals) == 1
    ref = float((x.double() ** 2).sum())
    assert abs(vals.pop() - ref) < 1e-5 * ref
    k.sumsq(x[:0], out, scratch)
    assert float(out) == 0.0
    k.sumsq(x[:8], out, scratch)                                      # one block
    assert abs(float(out) - float((x[:8].double() ** 2).sum())) < 1e-6


@pytest.mark.parametrize('B,N', [(3, 100), (1, 37), (2, 5), (64, 100)])
def test_relbias_tensor_core_mode_matches_fp64(B, N):
    """mode 1 of mmnas_relbias_fwd / _bwd (bf16 arm: warp-level mma.sync on register fragments, split-bf16 operands
    for r, plain bf16 operands for the gradient products) against the float64 formula and against the mode-0 kernels,
    8 heads, geometry input; pair counts that are not multiples of 32 / 16 exercise the dead-row handling."""
    k = K()
    R, h = 64, 8
    g4 = rnd(B, N, N, 4, seed=1)
    g4[0, N // 2:] = 0                    # zero-padded pairs, as the loader produces
    Wy, by = 0.5 * rnd(R, 4, seed=2), 0.1 * rnd(R, seed=3)
    Wr, br = 0.03 * rnd(h, R, seed=4), torch.ones(h, device=DEV)       # benign regime (see test_relbias_fwd_bwd)
    br[0] = -2.0
    go = rnd(B, h, N, N, seed=6)
    g4d, Wyd, byd, Wrd, brd = (t.double().requires_grad_(True) for t in (g4, Wy, by, Wr, br))
    e = torch.relu(g4d @ Wyd.t() + byd)
    r = torch.relu(e @ Wrd.t() + brd).permute(0, 3, 1, 2)
    ref = torch.log(torch.clamp(r, min=1e-6))
    ref.backward(go.double())
    bias = torch.empty(B, h, N, N, device=DEV)
    k.relbias_fwd(B, N, h, R, None, g4, Wy, by, Wr, br, bias, mode=1)
    safe = (r.detach() > 1e-4) | (r.detach() == 0)
    # split operands: r carries an absolute error ~1e-5 of its largest term, which the logarithm turns into a relative
    # one; at r = 1e-4 (the edge of `safe`) that is a few per cent of one logit unit, 4.5e-5 of the -13.8 range
    assert normwise(bias[safe], ref[safe]) < 2e-4
    dWr, dbr, dWy, dby = torch.zeros_like(Wr), torch.zeros_like(br), torch.zeros_like(Wy), torch.zeros_like(by)
    k.relbias_bwd(B, N, h, R, None, g4, Wy, by, Wr, br, go, None, dWy, dby, dWr, dbr, mode=1)
    tol = 1e-2                                             # bf16 operands in the gradient products
    assert normwise(dWr, Wrd.grad) < tol
    assert normwise(dbr, brd.grad) < tol
    assert normwise(dWy, Wyd.grad) < tol
    assert normwise(dby, byd.grad) < tol
    bias0 = torch.empty_like(bias)
    k.relbias_fwd(B, N, h, R, None, g4, Wy, by, Wr, br, bias0, mode=0)
    assert normwise(bias[safe], bias0[safe]) < 2e-4


def test_gemm_bf16_random_sweep():
    """Seeded sweep over sizes, operand layouts, tilings (single CTA 128 / 256 wide, CTA pair), split-K, output types,
    pitched outputs and epilogues of mmnas_gemm_bf16 against float64: ragged M, K not a multiple of 64 (TMA zero fill),
    N = 32 .. 2048, one-row problems."""
    import os
    import random
    k = K()
    rng = random.Random(1234)
    prev = {v: os.environ.get(v) for v in ('MMNAS_GEMM_PAIR', 'MMNAS_GEMM_BN')}
    try:
        for case in range(60):
            M = rng.choice([1, 7, 64, 127, 128, 129, 300, 896, 1000, 2049])
            N = rng.choice([32, 64, 96, 128, 256, 320, 512, 768, 2048])
            Kd = rng.choice([8, 16, 40, 64, 72, 200, 512, 1000])
            a_mn, b_mn = rng.randint(0, 1), rng.randint(0, 1)
            mode = rng.choice(['bf16', 'f32', 'f32_acc', 'splitk', 'bias_relu', 'aux', 'pitched'])
            for var, val in (('MMNAS_GEMM_PAIR', rng.choice([None, '0', '1'])), ('MMNAS_GEMM_BN', rng.choice([None, '128', '256']))):
                if val is None:
                    os.environ.pop(var, None)
                else:
                    os.environ[var] = val
            Kp, Mp = (Kd + 7) // 8 * 8, (M + 7) // 8 * 8
            A = rnd(Kp if a_mn else M, Mp if a_mn else Kp, seed=3 * case + 1, dtype=torch.bfloat16)
            B = rnd(Kp if b_mn else N, N if b_mn else Kp, seed=3 * case + 2, dtype=torch.bfloat16)
            Ad = (A[:Kd, :M].t() if a_mn else A[:, :Kd]).double()
            Bd = (B[:Kd] if b_mn else B[:, :Kd].t()).double()
            ref = Ad @ Bd
            tag = (case, M, N, Kd, a_mn, b_mn, mode, os.environ.get('MMNAS_GEMM_PAIR'), os.environ.get('MMNAS_GEMM_BN'))
            args = (M, N, Kd, A, A.stride(0), a_mn, B, B.stride(0), b_mn)
            if mode == 'bf16':
                C = torch.full((M, N), float('nan'), device=DEV, dtype=torch.bfloat16)
                k.gemm_bf16(*args, C, N)
                assert normwise(C, ref) < 6e-3, tag
            elif mode == 'f32':
                C = torch.full((M, N), float('nan'), device=DEV)
                k.gemm_bf16(*args, C, N)
                assert normwise(C, ref) < 1e-5, tag
            elif mode == 'f32_acc':
                C = rnd(M, N, seed=3 * case + 3)
                want = C.double() + ref
                k.gemm_bf16(*args, C, N, accumulate=True)
                assert normwise(C, want) < 1e-5, tag
            elif mode == 'splitk':
                C = torch.zeros(M, N, device=DEV)
                k.gemm_bf16(*args, C, N, split_k=rng.choice([2, 3, 5]))
                assert normwise(C, ref) < 1e-5, tag
            elif mode == 'bias_relu':
                bias = rnd(N, seed=3 * case + 3)
                C = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
                k.gemm_bf16(*args, C, N, bias=bias, relu=True)
                assert normwise(C, torch.relu(ref + bias.double())) < 6e-3, tag
            elif mode == 'aux':
                aux = rnd(M, N, seed=3 * case + 3, dtype=torch.bfloat16)
                C = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
                k.gemm_bf16(*args, C, N, aux=aux, ld_aux=N, aux_scale=1.25)
                assert normwise(C, torch.where(aux.double() > 0, ref * 1.25, torch.zeros_like(ref))) < 6e-3, tag
            else:
                big = torch.zeros(M, N + 64, device=DEV, dtype=torch.bfloat16)
                k.gemm_bf16(*args, big[:, 32:32 + N], N + 64)
                assert normwise(big[:, 32:32 + N], ref) < 6e-3, tag
                assert float(big[:, :32].abs().max()) == 0 and float(big[:, 32 + N:].abs().max()) == 0, tag
    finally:
        for var, val in prev.items():
            if val is None:
                os.environ.pop(var, None)
            else:
                os.environ[var] = val


def test_box_geometry_matches_reference_golden():
    """mmnas_box_geometry vs relation_embedding of the unmodified reference (load_data_vqa.py:7-33; golden geometry.npz,
    incl. two identical boxes that hit the 1e-3 clamp) and the loader's zero padding (:236-239)."""
    from mmnas_b200 import kernels as K
    from tests.util import load_golden
    r = load_golden('geometry.npz')
    boxes, rel = r['boxes'], r['rel']
    n = boxes.shape[0]
    g = K.box_geometry(boxes.view(1, n, 4).to(DEV), None)
    assert normwise(g[0], rel) < 2e-6
    N = 12                                                     # padded to 12 regions, two samples, ragged
    bx = torch.zeros(2, N, 4)
    bx[0, :n], bx[1, :5] = boxes, boxes[:5]
    pad = torch.ones(2, N, dtype=torch.uint8)
    pad[0, :n], pad[1, :5] = 0, 0
    g = K.box_geometry(bx.to(DEV), pad.to(DEV)).cpu()
    ref = torch.zeros(2, N, N, 4)
    ref[0, :n, :n], ref[1, :5, :5] = rel, rel[:5, :5]
    assert normwise(g, ref) < 2e-6
    assert torch.equal(g[0, n:], ref[0, n:]) and torch.equal(g[1, :, 5:], ref[1, :, 5:])      # padding exactly zero


@pytest.mark.parametrize('M,N,K', [(6400, 512, 512), (6400, 512, 2048), (896, 512, 2048), (324, 512, 512), (6400, 256, 256),
                                   (896, 256, 1024), (200, 256, 256)])
@pytest.mark.parametrize('variant', ['plain', 'bias_residual_dropout'])
def test_fused_gemm_layernorm_matches_the_two_kernel_tail(M, N, K, variant):
    """mmnas_gemm_ln_bf16 (projection + residual + dropout + LayerNorm in one tcgen05 cluster kernel, row statistics over
    distributed shared memory) against mmnas_gemm_bf16 followed by mmnas_ln_residual_fwd on the same operands: z, out,
    its bf16 copy, mean and sigma; same dropout stream (identical masks)."""
    from mmnas_b200 import kernels as K_
    import mmnas_b200
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).to(torch.bfloat16)
    gamma = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(N, generator=g)).to(DEV)
    full = variant != 'plain'
    bias = (0.5 * torch.randn(N, generator=g)).to(DEV) if full else None
    x = (torch.randn(M, N, generator=g) + 0.3).to(DEV) if full else None
    mmnas_b200.manual_seed(3)
    drop = K_.Drop(mmnas_b200.runtime.rng_state(DEV), 12345, 0.1) if full else K_.NO_DROP
    # reference: the two-kernel tail
    z_ref = torch.empty(M, N, device=DEV)
    K_.gemm_bf16(M, N, K, A, K, 0, W, K, 0, z_ref, N, bias=bias)
    out_ref, out16_ref = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    mean_ref, sigma_ref = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    K_.ln_residual_fwd(M, N, x, z_ref, gamma, beta, 1e-6, out_ref, out16_ref, mean_ref, sigma_ref, drop)
    # fused
    z, out = torch.full((M, N), float('nan'), device=DEV), torch.full((M, N), float('nan'), device=DEV)
    out16 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    mean, sigma = torch.zeros(M, device=DEV), torch.zeros(M, device=DEV)
    K_.gemm_ln_bf16(M, N, K, A, K, W, K, bias, x, gamma, beta, 1e-6, z, out, out16, mean, sigma, drop)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all() and torch.isfinite(z).all()
    assert normwise(z, z_ref) < 1e-6
    if full:
        assert torch.equal(z == x, z_ref == x)                 # the same elements were dropped
    assert normwise(mean, mean_ref, 1e-3) < 1e-5 and normwise(sigma, sigma_ref) < 1e-6
    assert normwise(out, out_ref) < 2e-6
    assert normwise(out16.float(), out16_ref.float()) < 8e-3   # one bf16 ulp where the fp32 values straddle a rounding boundary
    assert (out16 != out16_ref).float().mean() < 1e-3


# ------------------------------------------------------------------------------------------ text stem: LSTM
@pytest.mark.parametrize('B,T,E,H', [(64, 14, 300, 512), (192, 50, 300, 512), (64, 14, 300, 256), (5, 3, 40, 256), (33, 7, 300, 512), (130, 5, 64, 512), (8, 1, 300, 512)])
def test_native_lstm_matches_torch_lstm(B, T, E, H):
    """functional.LSTMFn (one GEMM for the input projection, one persistent cooperative kernel for the recurrence per
    direction, csrc/lstm.cu) against torch.nn.LSTM in float64: output sequence and every gradient (embedded input,
    W_ih, W_hh, both biases) within the bf16 arm's tolerance (full_vqa.py:68-74,94-95)."""
    import mmnas_b200
    from mmnas_b200.functional import LSTMFn
    torch.manual_seed(7)
    ref = torch.nn.LSTM(E, H, num_layers=1, batch_first=True).double()
    emb = torch.randn(B, T, E, dtype=torch.float64) * 0.5
    emb.requires_grad_(True)
    go = torch.randn(B, T, H, dtype=torch.float64)
    out_ref, _ = ref(emb)
    (out_ref * go).sum().backward()
    params = [p.detach().float().to(DEV).requires_grad_(True) for p in (ref.weight_ih_l0, ref.weight_hh_l0, ref.bias_ih_l0, ref.bias_hh_l0)]
    e32 = emb.detach().float().to(DEV).requires_grad_(True)
    with mmnas_b200.precision('bf16'):
        out = LSTMFn.apply(e32, *params)
        (out * go.float().to(DEV)).sum().backward()
    torch.cuda.synchronize()
    assert normwise(out, out_ref) < 1e-2
    assert normwise(e32.grad, emb.grad) < 2e-2
    for p, q in zip(params, (ref.weight_ih_l0, ref.weight_hh_l0, ref.bias_ih_l0, ref.bias_hh_l0)):
        assert normwise(p.grad, q.grad) < 2e-2
    # same inputs -> same bits (no atomics on the recurrent path)
    with mmnas_b200.precision('bf16'), torch.no_grad():
        out2 = LSTMFn.apply(e32, *params)
    assert torch.equal(out.detach(), out2)


def test_scalar_log_returns_pushed_values_in_order():
    """engine.ScalarLog: non-blocking device->host reads of per-step scalars, popped in push order."""
    from mmnas_b200.engine import ScalarLog
    log = ScalarLog(DEV, depth=3)
    src = torch.zeros((), device=DEV)
    got = []
    for i in range(7):
        src.fill_(float(i) + 0.5)            # same tensor every step, like a graph's static loss
        log.push(src)
        if len(log) > 1:
            got.append(log.pop())
    while len(log):
        got.append(log.pop())
    assert got == [i + 0.5 for i in range(7)]
    with pytest.raises(RuntimeError):
        log.pop()
