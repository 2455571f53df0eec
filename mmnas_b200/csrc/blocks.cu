// Block-level entry points (ABI v7): one C call enqueues the whole kernel sequence of a candidate block.
//   mmnas_mha_ln_{fwd,bwd}      SelfAtt / GuidedAtt   (reference modules.py:248-271, :301-325 over MHAtt :158-199)
//   mmnas_rel_mha_ln_{fwd,bwd}  RelSelfAtt            (modules.py:274-298 over RelMHAtt :202-245)
//   mmnas_ffn_ln_{fwd,bwd}      FeedForward           (modules.py:328-362 over MLP :34-41 / FC :13-31)
// Host code only: every step is one of the kernels of this library (tcgen05 GEMM / attention in the bf16 arm, FFMA
// kernels in the fp32 arm, relation bias, residual + LayerNorm, column sums), launched through the same extern "C"
// functions a caller could invoke one by one.  What moves here from the Python host is the composition: workspace
// carving, the order of the launches, the split-K policy of the weight gradients and the side-stream fork / join that
// lets the weight-gradient GEMMs fill SMs the dgrad / attention chain leaves idle.  A block forward or backward is then
// ONE foreign call instead of 5-9 (forward) / 9-14 (backward), which is what bounds the eager search step.
#include <atomic>
#include <cstdlib>
#include "common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

#define RC(call)                 \
  do {                           \
    int rc__ = (call);           \
    if (rc__ != MMNAS_OK) return rc__; \
  } while (0)

constexpr int HEAD = 64;

// Fused projection + residual + LayerNorm (gemm_ln.cu) wherever it applies; MMNAS_FUSE_LN=0 keeps the two-kernel tail
// (A/B measurements only).
bool fuse_ln() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MMNAS_FUSE_LN"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}

// branch GEMM + block tail: z = x + dropout(A W^T + bias), out = LN(z).  One launch when the fused kernel takes the
// shape, else the projection into z followed by the residual / LayerNorm kernel.
int tail16(int M, int H, int K, const void* A, const void* w16, const float* bias, const float* x, const float* ln_a,
           const float* ln_b, float eps, float* z, float* out, void* out16, float* mean, float* sigma,
           const unsigned long long* rng, unsigned long long salt, float p, mmnas_stream s) {
  // The fused kernel owns a 256-row block per cluster (2 CTAs at H = 256, 4 at H = 512) and runs one block per cluster:
  // it wins when the row blocks fill most of one wave of clusters (measured on B200, scripts/bench_gemm_ln.py:
  // 6400 x 512 x 512 17.1 vs 21.2 us, 6400 x 512 x 2048 25.5 vs 30.1 us) and loses on the short text stream (896 rows:
  // 13.7 vs 11.1 us, its epilogue is not overlapped with anything) or when the blocks spill into a second wave.
  const int row_blocks = (M + 255) / 256, per_wave = H == 256 ? 74 : 37;       // clusters of 2 / 4 CTAs on 148 SMs
  if (fuse_ln() && ln_a && out16 && (H == 256 || H == 512) && row_blocks <= per_wave && 2 * row_blocks > per_wave) {
    const int rc = mmnas_gemm_ln_bf16(M, H, K, A, K, w16, K, bias, x, ln_a, ln_b, eps, z, out, out16, mean, sigma, rng, salt, p, s);
    if (rc != MMNAS_ERR_UNSUPPORTED) return rc;
  }
  int rc = mmnas_gemm_bf16(M, H, K, A, K, 0, w16, K, 0, z, H, 0, bias, 0, 0, nullptr, 0, 1.f, 1, nullptr, 0, 0.f, s);
  if (rc != MMNAS_OK) return rc;
  return mmnas_ln_residual_fwd(M, H, x, z, ln_a, ln_b, eps, out, out16, mean, sigma, rng, salt, p, s);
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Split the token reduction of a weight-gradient GEMM so that (#output tiles x splits) ~ one wave of the GPU, keeping at
// least 4 k-blocks of 64 per split.  Large outputs with a long reduction go to the CTA-pair kernel (256 x 256 tiles, 74
// clusters; gemm_tc.cu switches to it at >= 48 units), everything else to 128 x 128 tiles on 148 SMs — fitted on B200.
inline int split_k(int M, int N, int K) {
  const int kb = (K + 63) / 64;
  const int cap = kb >= 4 ? kb / 4 : 1;
  if (N % 256 == 0 && M >= 256 && K >= 2048) {
    const int units = ((M + 255) / 256) * (N / 256);
    if (units >= 8 && units <= 74) { const int s = 74 / units < cap ? 74 / units : cap; return s < 1 ? 1 : s; }
  }
  const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
  const int w = tiles <= 148 ? 148 / tiles : 1;
  const int s = w < cap ? w : cap;
  return s < 1 ? 1 : s;
}

// ---- fork / join of the weight-gradient work on the side stream -------------------------------------------------
// Events come from a per-device ring of timing-disabled events (created on first use, reused round-robin; a wait
// captures the record that preceded it, so reuse is safe).  Inside a CUDA graph capture the record / wait pairs become
// the edges of a parallel branch.
constexpr int EV_RING = 256, MAX_DEV = 16;
cudaEvent_t g_events[MAX_DEV][EV_RING];
std::atomic<unsigned> g_ev_next[MAX_DEV];
std::atomic<int> g_ev_ready[MAX_DEV];

int next_event(cudaEvent_t* ev) {
  int dev = 0;
  MMNAS_CUDA(cudaGetDevice(&dev));
  MMNAS_CHECK_ARG(dev >= 0 && dev < MAX_DEV, "block: device index out of range");
  if (g_ev_ready[dev].load(std::memory_order_acquire) != 2) {
    int expect = 0;
    if (g_ev_ready[dev].compare_exchange_strong(expect, 1)) {
      for (int i = 0; i < EV_RING; ++i) MMNAS_CUDA(cudaEventCreateWithFlags(&g_events[dev][i], cudaEventDisableTiming));
      g_ev_ready[dev].store(2, std::memory_order_release);
    } else {
      while (g_ev_ready[dev].load(std::memory_order_acquire) != 2) {}
    }
  }
  *ev = g_events[dev][g_ev_next[dev].fetch_add(1) % EV_RING];
  return MMNAS_OK;
}

struct Fork {
  cudaStream_t main, side;
  bool enabled, used;
  Fork(mmnas_stream m, mmnas_stream s, bool on) : main((cudaStream_t)m), side((cudaStream_t)s), enabled(on && s != nullptr && s != m), used(false) {}
  // stream for work that may run concurrently with whatever is enqueued on `main` AFTER this call
  int begin(mmnas_stream* out) {
    if (!enabled) { *out = (mmnas_stream)main; return MMNAS_OK; }
    cudaEvent_t ev;
    RC(next_event(&ev));
    MMNAS_CUDA(cudaEventRecord(ev, main));
    MMNAS_CUDA(cudaStreamWaitEvent(side, ev, 0));
    used = true;
    *out = (mmnas_stream)side;
    return MMNAS_OK;
  }
  int join() {
    if (!enabled || !used) return MMNAS_OK;
    cudaEvent_t ev;
    RC(next_event(&ev));
    MMNAS_CUDA(cudaEventRecord(ev, side));
    MMNAS_CUDA(cudaStreamWaitEvent(main, ev, 0));
    used = false;
    return MMNAS_OK;
  }
};

inline int zero_f32(float* p, size_t n, mmnas_stream s) {
  MMNAS_CUDA(cudaMemsetAsync(p, 0, n * sizeof(float), (cudaStream_t)s));
  return MMNAS_OK;
}

// thin spellings of the primitive entry points with the defaults the blocks use
inline int gemm16(int M, int N, int K, const void* A, long lda, int a_mn, const void* B, long ldb, int b_mn, void* C, long ldc,
                  int out_bf16, const float* bias, int relu, int accumulate, int split, mmnas_stream s,
                  const void* aux = nullptr, long ld_aux = 0, float aux_scale = 1.f, const unsigned long long* rng = nullptr,
                  unsigned long long salt = 0, float p = 0.f) {
  return mmnas_gemm_bf16(M, N, K, A, lda, a_mn, B, ldb, b_mn, C, ldc, out_bf16, bias, relu, accumulate, aux, ld_aux, aux_scale,
                         split, rng, salt, p, s);
}
inline int gemm32(int M, int N, int K, const float* A, long a_rs, long a_cs, const float* B, long b_rs, long b_cs, float* C,
                  long ldc, const float* bias, int epilogue, int accumulate, mmnas_stream s, const float* aux = nullptr,
                  long ld_aux = 0, float aux_scale = 1.f, const unsigned long long* rng = nullptr, unsigned long long salt = 0,
                  float p = 0.f) {
  return mmnas_gemm_f32(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, epilogue, accumulate, aux, ld_aux, aux_scale, rng,
                        salt, p, s);
}

// ================================================================================================ attention blocks
struct AttLayout {
  size_t x16, kv16, qkv, kvb, bias, atted, z, mean, sigma, total;     // forward workspace (saved for backward)
  size_t dbranch, datt, dqkv, dkvb, dbias, btotal;                    // backward scratch
};

int att_check(const mmnas_att_block* d, bool rel, bool backward) {
  MMNAS_CHECK_ARG(d, "att block: null descriptor");
  MMNAS_CHECK_ARG(d->precision == 0 || d->precision == 1, "att block: precision must be 0 (fp32) or 1 (bf16)");
  MMNAS_CHECK_ARG(d->B >= 0 && d->Nq >= 1 && d->Nk >= 1 && d->H >= 4 && d->I >= HEAD && d->I % HEAD == 0,
                  "att block: bad sizes (inner width must be a multiple of the head dim 64)");
  MMNAS_CHECK_ARG(d->guided || d->Nk == d->Nq, "att block: self-attention needs Nk == Nq");
  MMNAS_CHECK_ARG((d->ln_a == nullptr) == (d->ln_b == nullptr), "att block: LayerNorm a_2 / b_2 must both be given");
  if (rel) {
    MMNAS_CHECK_ARG(d->R == 64 && !d->guided, "rel_mha_ln: relation attention is self-attention with R == 64");
    MMNAS_CHECK_ARG((d->rel != nullptr) != (d->g4 != nullptr), "rel_mha_ln: exactly one of rel / g4 must be given");
    MMNAS_CHECK_ARG(d->Wr && d->br && (!d->g4 || (d->Wy && d->by)), "rel_mha_ln: relation parameters missing");
  } else {
    MMNAS_CHECK_ARG(d->R == 0 && !d->rel && !d->g4, "mha_ln: relation inputs given (use mmnas_rel_mha_ln_*)");
  }
  if (d->B == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(d->x && d->workspace && d->Wq && d->Wk && d->Wv && d->Wm, "att block: null input / parameter / workspace");
  MMNAS_CHECK_ARG(!d->guided || d->kv, "mha_ln: guided attention needs kv");
  MMNAS_CHECK_ARG(((uintptr_t)d->workspace & 255) == 0, "att block: workspace must be 256-byte aligned");
  if (d->precision == 1)
    MMNAS_CHECK_ARG(d->w16_a && d->w16_m && (!d->guided || d->w16_b), "att block: bf16 weight copies missing");
  if (!backward) {
    MMNAS_CHECK_ARG(d->out, "att block: null output");
  } else {
    MMNAS_CHECK_ARG(d->dout && d->dx && d->bwd_workspace && d->dWq && d->dWk && d->dWv && d->dWm,
                    "att block bwd: null gradient buffer / workspace");
    MMNAS_CHECK_ARG(((uintptr_t)d->bwd_workspace & 255) == 0, "att block: bwd_workspace must be 256-byte aligned");
    MMNAS_CHECK_ARG(!d->guided || d->dkv, "mha_ln bwd: guided attention needs dkv");
    MMNAS_CHECK_ARG(!d->ln_a || (d->dln_a && d->dln_b), "att block bwd: LayerNorm gradient buffers missing");
    if (rel) {
      MMNAS_CHECK_ARG(d->dWr && d->dbr && (!d->g4 || (d->dWy && d->dby)) && (!d->rel || d->drel),
                      "rel_mha_ln bwd: relation gradient buffers missing");
    }
  }
  return MMNAS_OK;
}

AttLayout att_layout(const mmnas_att_block* d) {
  AttLayout L = {};
  const size_t Mq = (size_t)d->B * d->Nq, Mk = (size_t)d->B * d->Nk, H = d->H, I = d->I, heads = d->I / HEAD;
  const bool bf = d->precision == 1;
  const size_t es = bf ? 2 : 4;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align256(o + bytes); return at; };
  L.x16 = take(bf ? Mq * H * 2 : 0);
  L.kv16 = take(bf && d->guided ? Mk * H * 2 : 0);
  L.qkv = take(Mq * (d->guided ? I : 3 * I) * es);
  L.kvb = take(d->guided ? Mk * 2 * I * es : 0);
  L.bias = take(d->R ? (size_t)d->B * heads * d->Nq * d->Nk * 4 : 0);
  L.atted = take(Mq * I * es);
  L.z = take(Mq * H * 4);
  L.mean = take(Mq * 4);
  L.sigma = take(Mq * 4);
  L.total = o < 256 ? 256 : o;
  o = 0;
  L.dbranch = take(Mq * H * es);
  L.datt = take(Mq * I * es);
  L.dqkv = take(Mq * (d->guided ? I : 3 * I) * es);
  L.dkvb = take(d->guided ? Mk * 2 * I * es : 0);
  L.dbias = take(d->R ? (size_t)d->B * heads * d->Nq * d->Nk * 4 : 0);
  L.btotal = o < 256 ? 256 : o;
  return L;
}

struct AttPtrs {   // operand views into the fused projection buffers (registration order v, k, q)
  char *q, *k, *v; long ldq, ldk, ldv;
};
AttPtrs att_views(const mmnas_att_block* d, char* qkv, char* kvb) {
  const size_t es = d->precision == 1 ? 2 : 4, I = d->I;
  AttPtrs p;
  if (!d->guided) {
    p.v = qkv; p.k = qkv + I * es; p.q = qkv + 2 * I * es;
    p.ldq = p.ldk = p.ldv = 3 * (long)I;
  } else {
    p.q = qkv; p.ldq = (long)I;
    p.v = kvb; p.k = kvb + I * es; p.ldk = p.ldv = 2 * (long)I;
  }
  return p;
}

int att_fwd(const mmnas_att_block* d, bool rel) {
  RC(att_check(d, rel, false));
  if (d->B == 0) return MMNAS_OK;
  const AttLayout L = att_layout(d);
  char* ws = (char*)d->workspace;
  const int B = d->B, Nq = d->Nq, Nk = d->Nk, H = d->H, I = d->I, heads = I / HEAD;
  const int Mq = B * Nq, Mk = B * Nk;
  const bool bf = d->precision == 1;
  mmnas_stream s = d->stream;
  const unsigned long long* rng = d->rng_state;
  // --- bf16 operand copies of the inputs (the previous block's LayerNorm kernel normally provides them)
  const void* x16 = d->x16;
  const void* kv16 = d->kv16;
  if (bf && !x16) { RC(mmnas_cast_f32_to_bf16(d->x, ws + L.x16, (long)Mq * H, s)); x16 = ws + L.x16; }
  if (bf && d->guided && !kv16) { RC(mmnas_cast_f32_to_bf16(d->kv, ws + L.kv16, (long)Mk * H, s)); kv16 = ws + L.kv16; }
  // --- projections into the fused buffers
  char* qkv = ws + L.qkv;
  char* kvb = ws + L.kvb;
  const AttPtrs P = att_views(d, qkv, kvb);
  if (!d->guided) {
    if (bf) {
      RC(gemm16(Mq, 3 * I, H, x16, H, 0, d->w16_a, H, 0, qkv, 3 * I, 1, nullptr, 0, 0, 1, s));
    } else {
      RC(gemm32(Mq, I, H, d->x, H, 1, d->Wq, 1, H, (float*)P.q, 3 * I, nullptr, 0, 0, s));
      RC(gemm32(Mq, I, H, d->x, H, 1, d->Wk, 1, H, (float*)P.k, 3 * I, nullptr, 0, 0, s));
      RC(gemm32(Mq, I, H, d->x, H, 1, d->Wv, 1, H, (float*)P.v, 3 * I, nullptr, 0, 0, s));
    }
  } else {
    if (bf) {
      RC(gemm16(Mq, I, H, x16, H, 0, d->w16_a, H, 0, qkv, I, 1, nullptr, 0, 0, 1, s));
      RC(gemm16(Mk, 2 * I, H, kv16, H, 0, d->w16_b, H, 0, kvb, 2 * I, 1, nullptr, 0, 0, 1, s));
    } else {
      RC(gemm32(Mq, I, H, d->x, H, 1, d->Wq, 1, H, (float*)P.q, I, nullptr, 0, 0, s));
      RC(gemm32(Mk, I, H, d->kv, H, 1, d->Wk, 1, H, (float*)P.k, 2 * I, nullptr, 0, 0, s));
      RC(gemm32(Mk, I, H, d->kv, H, 1, d->Wv, 1, H, (float*)P.v, 2 * I, nullptr, 0, 0, s));
    }
  }
  // --- RSA logit bias from the geometry path
  float* bias = nullptr;
  if (d->R) {
    bias = (float*)(ws + L.bias);
    if (d->g4) RC(mmnas_relbias_fwd(bf ? 1 : 0, B, Nq, heads, d->R, nullptr, d->g4, d->Wy, d->by, d->Wr, d->br, bias, s));
    else RC(mmnas_relbias_fwd(0, B, Nq, heads, d->R, d->rel, nullptr, nullptr, nullptr, d->Wr, d->br, bias, s));
  }
  // --- attention core, merged-head output
  void* atted = ws + L.atted;
  RC(mmnas_attn_fwd(bf ? 1 : 0, B, heads, Nq, Nk, HEAD, P.q, P.ldq, P.k, P.ldk, P.v, P.ldv, d->kmask, bias, atted, I, 0.125f,
                    d->p_att > 0.f ? rng : nullptr, d->salt_att, d->p_att, s));
  // --- merge projection, residual, LayerNorm (z = x + dropout(branch) overwrites the branch buffer)
  float* z = (float*)(ws + L.z);
  const unsigned long long* rng_out = d->p_out > 0.f ? rng : nullptr;
  if (bf)
    return tail16(Mq, H, I, atted, d->w16_m, nullptr, d->residual ? d->x : nullptr, d->ln_a, d->ln_b, d->eps, z, d->out, d->out16,
                  (float*)(ws + L.mean), (float*)(ws + L.sigma), rng_out, d->salt_out, d->p_out, s);
  RC(gemm32(Mq, H, I, (const float*)atted, I, 1, d->Wm, 1, I, z, H, nullptr, 0, 0, s));
  RC(mmnas_ln_residual_fwd(Mq, H, d->residual ? d->x : nullptr, z, d->ln_a, d->ln_b, d->eps, d->out, d->out16,
                           (float*)(ws + L.mean), (float*)(ws + L.sigma), rng_out, d->salt_out, d->p_out, s));
  return MMNAS_OK;
}

// C (+)= A^T Bm over `tokens` rows: A [tokens, M] (pitch lda), Bm [tokens, N] (pitch ldb), C [M, N] fp32
int wgrad(const mmnas_att_block* d, Fork& fork, int M, int N, int tokens, const void* A, long lda, const void* Bm, long ldb,
          const float* a32, const float* b32, float* C, int accumulate) {
  if (d->precision == 1) {
    const int sk = split_k(M, N, tokens);
    if (sk > 1 && !accumulate) RC(zero_f32(C, (size_t)M * N, d->stream));
    mmnas_stream ws;
    RC(fork.begin(&ws));
    return gemm16(M, N, tokens, A, lda, 1, Bm, ldb, 1, C, N, 0, nullptr, 0, accumulate && sk == 1, sk, ws);
  }
  return gemm32(M, N, tokens, a32, 1, lda, b32, ldb, 1, C, N, nullptr, 0, accumulate, d->stream);
}

int att_bwd(const mmnas_att_block* d, bool rel) {
  RC(att_check(d, rel, true));
  if (d->B == 0) return MMNAS_OK;
  const AttLayout L = att_layout(d);
  char* ws = (char*)d->workspace;
  char* bw = (char*)d->bwd_workspace;
  const int B = d->B, Nq = d->Nq, Nk = d->Nk, H = d->H, I = d->I, heads = I / HEAD;
  const int Mq = B * Nq, Mk = B * Nk;
  const bool bf = d->precision == 1;
  const int acc = d->accumulate_grads;
  mmnas_stream s = d->stream;
  const unsigned long long* rng = d->rng_state;
  const bool norm = d->ln_a != nullptr;
  const bool drop_out = rng && d->p_out > 0.f;
  Fork fork(d->stream, d->side_stream, bf);

  const void* x16 = d->x16 ? d->x16 : (const void*)(ws + L.x16);       // the forward cast it there when it was not given
  const void* kv16 = d->kv16 ? d->kv16 : (const void*)(ws + L.kv16);
  char* qkv = ws + L.qkv;
  char* kvb = ws + L.kvb;
  const AttPtrs P = att_views(d, qkv, kvb);
  float* bias = d->R ? (float*)(ws + L.bias) : nullptr;
  void* atted = ws + L.atted;
  float* z = (float*)(ws + L.z);

  // --- LayerNorm + residual + output-dropout backward.  dx receives dL/dz (= the residual path's dL/dx); the branch
  // gradient gets its own buffer when it differs from it (dropout mask, bf16 operand type, no residual).
  const bool separate = bf || drop_out || !d->residual;
  void* dbranch = separate ? (void*)(bw + L.dbranch) : (void*)d->dx;
  if (norm && !acc) { RC(zero_f32(d->dln_a, H, s)); RC(zero_f32(d->dln_b, H, s)); }
  RC(mmnas_ln_residual_bwd(Mq, H, d->dout, z, (const float*)(ws + L.mean), (const float*)(ws + L.sigma), d->ln_a, d->eps,
                           d->residual ? d->dx : nullptr, separate ? dbranch : nullptr, separate && bf ? 1 : 0,
                           norm ? d->dln_a : nullptr, norm ? d->dln_b : nullptr, drop_out ? rng : nullptr, d->salt_out,
                           d->p_out, s));
  // --- merge projection backward
  void* datt = bw + L.datt;
  RC(wgrad(d, fork, H, I, Mq, dbranch, H, atted, I, (const float*)dbranch, (const float*)atted, d->dWm, acc));
  if (bf) RC(gemm16(Mq, I, H, dbranch, H, 0, d->w16_m, I, 1, datt, I, 1, nullptr, 0, 0, 1, s));
  else RC(gemm32(Mq, I, H, (const float*)dbranch, H, 1, d->Wm, I, 1, (float*)datt, I, nullptr, 0, 0, s));
  // --- attention core backward into the fused gradient buffers
  char* dqkv = bw + L.dqkv;
  char* dkvb = bw + L.dkvb;
  const AttPtrs G = att_views(d, dqkv, dkvb);
  float* dbias = d->R ? (float*)(bw + L.dbias) : nullptr;
  RC(mmnas_attn_bwd(bf ? 1 : 0, B, heads, Nq, Nk, HEAD, P.q, P.ldq, P.k, P.ldk, P.v, P.ldv, d->kmask, bias, atted, I, datt, I,
                    G.q, G.ldq, G.k, G.ldk, G.v, G.ldv, dbias, 0.125f, d->p_att > 0.f ? rng : nullptr, d->salt_att, d->p_att, s));
  // --- geometry-bias backward (the kernels accumulate with atomics)
  if (d->R) {
    if (!acc) { RC(zero_f32(d->dWr, (size_t)heads * d->R, s)); RC(zero_f32(d->dbr, heads, s)); }
    if (d->g4) {
      if (!d->accumulate_geometry) { RC(zero_f32(d->dWy, (size_t)d->R * 4, s)); RC(zero_f32(d->dby, d->R, s)); }
      RC(mmnas_relbias_bwd(bf ? 1 : 0, B, Nq, heads, d->R, nullptr, d->g4, d->Wy, d->by, d->Wr, d->br, dbias, nullptr, d->dWy,
                           d->dby, d->dWr, d->dbr, s));
    } else {
      RC(mmnas_relbias_bwd(0, B, Nq, heads, d->R, d->rel, nullptr, nullptr, nullptr, d->Wr, d->br, dbias, d->drel, nullptr,
                           nullptr, d->dWr, d->dbr, s));
    }
  }
  // --- projection backward: weight gradients (side stream) and input gradients
  const int acc_dx = d->residual ? 1 : 0;          // dx already holds dL/dz from the LayerNorm backward
  const size_t IH = (size_t)I * H;
  if (!d->guided) {
    if (d->dWk == d->dWv + IH && d->dWq == d->dWk + IH) {             // [dWv; dWk; dWq] contiguous: one GEMM
      RC(wgrad(d, fork, 3 * I, H, Mq, dqkv, 3 * I, x16, H, (const float*)dqkv, d->x, d->dWv, acc));
    } else {
      RC(wgrad(d, fork, I, H, Mq, G.v, 3 * I, x16, H, (const float*)G.v, d->x, d->dWv, acc));
      RC(wgrad(d, fork, I, H, Mq, G.k, 3 * I, x16, H, (const float*)G.k, d->x, d->dWk, acc));
      RC(wgrad(d, fork, I, H, Mq, G.q, 3 * I, x16, H, (const float*)G.q, d->x, d->dWq, acc));
    }
    if (bf) {
      RC(gemm16(Mq, H, 3 * I, dqkv, 3 * I, 0, d->w16_a, H, 1, d->dx, H, 0, nullptr, 0, acc_dx, 1, s));
    } else {
      RC(gemm32(Mq, H, I, (const float*)G.v, 3 * I, 1, d->Wv, H, 1, d->dx, H, nullptr, 0, acc_dx, s));
      RC(gemm32(Mq, H, I, (const float*)G.k, 3 * I, 1, d->Wk, H, 1, d->dx, H, nullptr, 0, 1, s));
      RC(gemm32(Mq, H, I, (const float*)G.q, 3 * I, 1, d->Wq, H, 1, d->dx, H, nullptr, 0, 1, s));
    }
  } else {
    RC(wgrad(d, fork, I, H, Mq, dqkv, I, x16, H, (const float*)dqkv, d->x, d->dWq, acc));
    if (d->dWk == d->dWv + IH) {
      RC(wgrad(d, fork, 2 * I, H, Mk, dkvb, 2 * I, kv16, H, (const float*)dkvb, d->kv, d->dWv, acc));
    } else {
      RC(wgrad(d, fork, I, H, Mk, G.v, 2 * I, kv16, H, (const float*)G.v, d->kv, d->dWv, acc));
      RC(wgrad(d, fork, I, H, Mk, G.k, 2 * I, kv16, H, (const float*)G.k, d->kv, d->dWk, acc));
    }
    if (bf) {
      RC(gemm16(Mq, H, I, dqkv, I, 0, d->w16_a, H, 1, d->dx, H, 0, nullptr, 0, acc_dx, 1, s));
      RC(gemm16(Mk, H, 2 * I, dkvb, 2 * I, 0, d->w16_b, H, 1, d->dkv, H, 0, nullptr, 0, d->accumulate_dkv ? 1 : 0, 1, s));
    } else {
      RC(gemm32(Mq, H, I, (const float*)G.q, I, 1, d->Wq, H, 1, d->dx, H, nullptr, 0, acc_dx, s));
      RC(gemm32(Mk, H, I, (const float*)G.v, 2 * I, 1, d->Wv, H, 1, d->dkv, H, nullptr, 0, d->accumulate_dkv ? 1 : 0, s));
      RC(gemm32(Mk, H, I, (const float*)G.k, 2 * I, 1, d->Wk, H, 1, d->dkv, H, nullptr, 0, 1, s));
    }
  }
  return fork.join();
}

// ================================================================================================ FeedForward
struct FfnLayout {
  size_t x16, h, z, mean, sigma, total;
  size_t dbranch, dh, btotal;
};

FfnLayout ffn_layout(const mmnas_ffn_block* d) {
  FfnLayout L = {};
  const size_t M = d->M, H = d->H, F = d->F;
  const bool bf = d->precision == 1;
  const size_t es = bf ? 2 : 4;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align256(o + bytes); return at; };
  L.x16 = take(bf ? M * H * 2 : 0);
  L.h = take(M * F * es);
  L.z = take(M * H * 4);
  L.mean = take(M * 4);
  L.sigma = take(M * 4);
  L.total = o < 256 ? 256 : o;
  o = 0;
  L.dbranch = take(M * H * es);
  L.dh = take(M * F * es);
  L.btotal = o < 256 ? 256 : o;
  return L;
}

int ffn_check(const mmnas_ffn_block* d, bool backward) {
  MMNAS_CHECK_ARG(d, "ffn_ln: null descriptor");
  MMNAS_CHECK_ARG(d->precision == 0 || d->precision == 1, "ffn_ln: precision must be 0 (fp32) or 1 (bf16)");
  MMNAS_CHECK_ARG(d->M >= 0 && d->H >= 4 && d->F >= 4, "ffn_ln: bad sizes");
  MMNAS_CHECK_ARG((d->ln_a == nullptr) == (d->ln_b == nullptr), "ffn_ln: LayerNorm a_2 / b_2 must both be given");
  if (d->M == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(d->x && d->W1 && d->b1 && d->W2 && d->b2 && d->workspace, "ffn_ln: null input / parameter / workspace");
  MMNAS_CHECK_ARG(((uintptr_t)d->workspace & 255) == 0, "ffn_ln: workspace must be 256-byte aligned");
  if (d->precision == 1) MMNAS_CHECK_ARG(d->w16_1 && d->w16_2, "ffn_ln: bf16 weight copies missing");
  if (!backward) {
    MMNAS_CHECK_ARG(d->out, "ffn_ln: null output");
  } else {
    MMNAS_CHECK_ARG(d->dout && d->dx && d->bwd_workspace && d->dW1 && d->db1 && d->dW2 && d->db2, "ffn_ln bwd: null gradient buffer");
    MMNAS_CHECK_ARG(((uintptr_t)d->bwd_workspace & 255) == 0, "ffn_ln: bwd_workspace must be 256-byte aligned");
    MMNAS_CHECK_ARG(!d->ln_a || (d->dln_a && d->dln_b), "ffn_ln bwd: LayerNorm gradient buffers missing");
  }
  return MMNAS_OK;
}

int ffn_fwd(const mmnas_ffn_block* d) {
  RC(ffn_check(d, false));
  if (d->M == 0) return MMNAS_OK;
  const FfnLayout L = ffn_layout(d);
  char* ws = (char*)d->workspace;
  const int M = d->M, H = d->H, F = d->F;
  const bool bf = d->precision == 1;
  mmnas_stream s = d->stream;
  const unsigned long long* rng = d->rng_state;
  const bool drop_mid = rng && d->p_mid > 0.f;
  void* h = ws + L.h;
  float* z = (float*)(ws + L.z);
  if (bf) {
    const void* x16 = d->x16;
    if (!x16) { RC(mmnas_cast_f32_to_bf16(d->x, ws + L.x16, (long)M * H, s)); x16 = ws + L.x16; }
    // bias, ReLU and the hidden dropout run in the epilogue of the first GEMM; the second adds its bias
    RC(gemm16(M, F, H, x16, H, 0, d->w16_1, H, 0, h, F, 1, d->b1, 1, 0, 1, s, nullptr, 0, 1.f, drop_mid ? rng : nullptr,
              d->salt_mid, drop_mid ? d->p_mid : 0.f));
    return tail16(M, H, F, h, d->w16_2, d->b2, d->residual ? d->x : nullptr, d->ln_a, d->ln_b, d->eps, z, d->out, d->out16,
                  (float*)(ws + L.mean), (float*)(ws + L.sigma), d->p_out > 0.f ? rng : nullptr, d->salt_out, d->p_out, s);
  }
  RC(gemm32(M, F, H, d->x, H, 1, d->W1, 1, H, (float*)h, F, d->b1, drop_mid ? 2 : 1, 0, s, nullptr, 0, 1.f,
            drop_mid ? rng : nullptr, d->salt_mid, drop_mid ? d->p_mid : 0.f));
  RC(gemm32(M, H, F, (const float*)h, F, 1, d->W2, 1, F, z, H, d->b2, 0, 0, s));
  RC(mmnas_ln_residual_fwd(M, H, d->residual ? d->x : nullptr, z, d->ln_a, d->ln_b, d->eps, d->out, d->out16,
                           (float*)(ws + L.mean), (float*)(ws + L.sigma), d->p_out > 0.f ? rng : nullptr, d->salt_out, d->p_out, s));
  return MMNAS_OK;
}

int ffn_bwd(const mmnas_ffn_block* d) {
  RC(ffn_check(d, true));
  if (d->M == 0) return MMNAS_OK;
  const FfnLayout L = ffn_layout(d);
  char* ws = (char*)d->workspace;
  char* bw = (char*)d->bwd_workspace;
  const int M = d->M, H = d->H, F = d->F;
  const bool bf = d->precision == 1;
  const int acc = d->accumulate_grads;
  mmnas_stream s = d->stream;
  const unsigned long long* rng = d->rng_state;
  const bool norm = d->ln_a != nullptr;
  const bool drop_out = rng && d->p_out > 0.f, drop_mid = rng && d->p_mid > 0.f;
  const void* h = ws + L.h;
  const float* z = (const float*)(ws + L.z);
  const bool separate = bf || drop_out || !d->residual;
  void* dbranch = separate ? (void*)(bw + L.dbranch) : (void*)d->dx;
  if (norm && !acc) { RC(zero_f32(d->dln_a, H, s)); RC(zero_f32(d->dln_b, H, s)); }
  RC(mmnas_ln_residual_bwd(M, H, d->dout, z, (const float*)(ws + L.mean), (const float*)(ws + L.sigma), d->ln_a, d->eps,
                           d->residual ? d->dx : nullptr, separate ? dbranch : nullptr, separate && bf ? 1 : 0,
                           norm ? d->dln_a : nullptr, norm ? d->dln_b : nullptr, drop_out ? rng : nullptr, d->salt_out,
                           d->p_out, s));
  const float keep_scale = drop_mid ? 1.f / (1.f - d->p_mid) : 1.f;
  const int acc_dx = d->residual ? 1 : 0;
  void* dh = bw + L.dh;
  if (bf) {
    const void* x16 = d->x16 ? d->x16 : (const void*)(ws + L.x16);
    Fork fork(d->stream, d->side_stream, true);
    const int sk2 = split_k(H, F, M), sk1 = split_k(F, H, M);
    if (sk2 > 1 && !acc) RC(zero_f32(d->dW2, (size_t)H * F, s));
    if (sk1 > 1 && !acc) RC(zero_f32(d->dW1, (size_t)F * H, s));
    mmnas_stream side;
    RC(fork.begin(&side));                    // db2, dW2 need only dbranch: they overlap the dh GEMM
    RC(mmnas_colsum(1, dbranch, M, H, H, d->db2, acc, side));
    RC(gemm16(H, F, M, dbranch, H, 1, h, F, 1, d->dW2, F, 0, nullptr, 0, acc && sk2 == 1, sk2, side));
    RC(gemm16(M, F, H, dbranch, H, 0, d->w16_2, F, 1, dh, F, 1, nullptr, 0, 0, 1, s, h, F, keep_scale));
    RC(fork.begin(&side));                    // db1, dW1 need dh: they overlap the dx GEMM
    RC(mmnas_colsum(1, dh, M, F, F, d->db1, acc, side));
    RC(gemm16(F, H, M, dh, F, 1, x16, H, 1, d->dW1, H, 0, nullptr, 0, acc && sk1 == 1, sk1, side));
    RC(gemm16(M, H, F, dh, F, 0, d->w16_1, H, 1, d->dx, H, 0, nullptr, 0, acc_dx, 1, s));
    return fork.join();
  }
  RC(mmnas_colsum(0, dbranch, M, H, H, d->db2, acc, s));
  RC(gemm32(H, F, M, (const float*)dbranch, 1, H, (const float*)h, F, 1, d->dW2, F, nullptr, 0, acc, s));
  RC(gemm32(M, F, H, (const float*)dbranch, H, 1, d->W2, F, 1, (float*)dh, F, nullptr, 3, 0, s, (const float*)h, F, keep_scale));
  RC(mmnas_colsum(0, dh, M, F, F, d->db1, acc, s));
  RC(gemm32(F, H, M, (const float*)dh, 1, F, d->x, H, 1, d->dW1, H, nullptr, 0, acc, s));
  RC(gemm32(M, H, F, (const float*)dh, F, 1, d->W1, H, 1, d->dx, H, nullptr, 0, acc_dx, s));
  return MMNAS_OK;
}

}  // namespace

extern "C" int mmnas_att_block_sizeof(void) { return (int)sizeof(mmnas_att_block); }
extern "C" int mmnas_ffn_block_sizeof(void) { return (int)sizeof(mmnas_ffn_block); }

extern "C" int mmnas_att_block_workspace(const mmnas_att_block* d, unsigned long long* fwd_bytes, unsigned long long* bwd_bytes) {
  MMNAS_CHECK_ARG(d && fwd_bytes && bwd_bytes, "att_block_workspace: null argument");
  MMNAS_CHECK_ARG(d->B >= 0 && d->Nq >= 1 && d->Nk >= 1 && d->H >= 4 && d->I >= HEAD && d->I % HEAD == 0, "att_block_workspace: bad sizes");
  const AttLayout L = att_layout(d);
  *fwd_bytes = L.total;
  *bwd_bytes = L.btotal;
  return MMNAS_OK;
}
extern "C" int mmnas_ffn_block_workspace(const mmnas_ffn_block* d, unsigned long long* fwd_bytes, unsigned long long* bwd_bytes) {
  MMNAS_CHECK_ARG(d && fwd_bytes && bwd_bytes, "ffn_block_workspace: null argument");
  MMNAS_CHECK_ARG(d->M >= 0 && d->H >= 4 && d->F >= 4, "ffn_block_workspace: bad sizes");
  const FfnLayout L = ffn_layout(d);
  *fwd_bytes = L.total;
  *bwd_bytes = L.btotal;
  return MMNAS_OK;
}
extern "C" int mmnas_mha_ln_fwd(const mmnas_att_block* d) { return att_fwd(d, false); }
extern "C" int mmnas_mha_ln_bwd(const mmnas_att_block* d) { return att_bwd(d, false); }
extern "C" int mmnas_rel_mha_ln_fwd(const mmnas_att_block* d) { return att_fwd(d, true); }
extern "C" int mmnas_rel_mha_ln_bwd(const mmnas_att_block* d) { return att_bwd(d, true); }
extern "C" int mmnas_ffn_ln_fwd(const mmnas_ffn_block* d) { return ffn_fwd(d); }
extern "C" int mmnas_ffn_ln_bwd(const mmnas_ffn_block* d) { return ffn_bwd(d); }
