// RSA geometry path -> attention-logit bias.
//   bias[b,h,i,j] = log(max(relu(W_r . e_ij + b_r)[h], 1e-6))            (modules.py:231,235)
//   e_ij = rel_embed[b,i,j,:]                                  (dense mode: the reference's tensor)
//        = relu(W_y . g_ij + b_y),  g_ij = 4-d box log-geometry (geometry mode: full_vqa.py:82,103
//          folded in, so the [B,N,N,REL_SIZE] tensor never exists in HBM)
// Backward accumulates dW_r, db_r (and dW_y, db_y in geometry mode; d rel_embed in dense mode) from
// dbias = dS of the attention backward.
//
// Design (v3).  Plain fp32 on the CUDA cores (both precision arms share it; the geometry path is too
// ill-conditioned for bf16 operands), organised so that the FP32 pipe, not instruction issue, is the limit:
//  * every multiply-add is a packed FFMA2 (fma.rn.f32x2, two IEEE fp32 FMAs per issue slot);
//  * the layer weights live in shared memory as ready-made pairs, staged by each CTA from the parameter tensors
//    (no constant-bank staging copies, so the entry points are stream-safe);
//  * forward: one thread per four pairs, FFMA2 packed over neighbouring rel channels (c, c+1);
//  * backward: 128-pair tiles.  Phase A (thread = pair, packed over channels) recomputes e and r and parks e
//    (transposed, padded) plus d pre_r / g (transposed) in shared memory.  Phase B (thread = channel c over half
//    of the tile, packed over neighbouring PAIRS) forms d e = W_r[:,c] . d pre_r with its own weight column in
//    registers and reduces dW_r[:,c], dW_y[c,:], db_y[c] with a two-level sum; one global atomic per output per
//    CTA at the end.  44 KB of shared memory per CTA, persistent grid of 4 CTAs per SM.
#include "common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int MAXH = 16;     // heads
constexpr int R = 64;        // REL_SIZE
constexpr int R2 = R / 2;    // channel pairs
constexpr int TILE = 128;    // pairs per backward tile == threads per CTA
constexpr int EP = TILE + 4; // padded pair pitch of the transposed tiles

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void unpk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float sum2(u64 v) {
  float lo, hi;
  unpk2(v, lo, hi);
  return lo + hi;
}

struct RelArgs {
  int B, N, heads;
  unsigned pairs, nn;
  const float* rel;     // dense: [pairs, R]
  const float* g4;      // geometry: [pairs, 4]
  const float *Wy, *by, *Wr, *br;
  float* bias;          // [B, heads, N, N]
  const float* dbias;
  float* drel;
  float *dWy, *dby, *dWr, *dbr;
};

// shared-memory weights, as channel pairs: Wy2[c2][k] = (Wy[2c2][k], Wy[2c2+1][k]), by2[c2], Wr2[c2][h] =
// (Wr[h][2c2], Wr[h][2c2+1]), then br[h]
template <int HEADS> constexpr int sw_floats() { return 2 * (R2 * 4 + R2 + R2 * HEADS) + MAXH; }

template <int HEADS, bool DENSE>
__device__ __forceinline__ void stage_weights(float* sw, const RelArgs& a, int tid, int nthr) {
  float2* Wy2 = reinterpret_cast<float2*>(sw);
  float2* by2 = Wy2 + R2 * 4;
  float2* Wr2 = by2 + R2;
  float* br = reinterpret_cast<float*>(Wr2 + R2 * HEADS);
  if (!DENSE) {
    for (int i = tid; i < R2 * 4; i += nthr) {
      const int c2 = i >> 2, k = i & 3;
      Wy2[i] = make_float2(__ldg(a.Wy + (2 * c2) * 4 + k), __ldg(a.Wy + (2 * c2 + 1) * 4 + k));
    }
    for (int i = tid; i < R2; i += nthr) by2[i] = make_float2(__ldg(a.by + 2 * i), __ldg(a.by + 2 * i + 1));
  }
  for (int i = tid; i < R2 * HEADS; i += nthr) {
    const int c2 = i / HEADS, h = i - c2 * HEADS;
    Wr2[i] = make_float2(__ldg(a.Wr + h * R + 2 * c2), __ldg(a.Wr + h * R + 2 * c2 + 1));
  }
  for (int i = tid; i < HEADS; i += nthr) br[i] = __ldg(a.br + i);
}

// e (channel pair c2) and r += W_r e for ONE pair; `gk` = the pair's geometry, each component duplicated
template <int HEADS>
__device__ __forceinline__ u64 geo_channel_pair(const u64* Wy2, const u64* by2, const u64 (&gk)[4], int c2) {
  u64 e2 = by2[c2];
#pragma unroll
  for (int k = 0; k < 4; ++k) e2 = ffma2(Wy2[c2 * 4 + k], gk[k], e2);
  float lo, hi;
  unpk2(e2, lo, hi);
  return pk2(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
}

// Forward: one thread per FPT pairs.  A broadcast LDS delivers 16 bytes to each of 32 lanes, i.e. costs four cycles of
// the SM's 128 B/clk shared-memory return path, so the weight pairs are amortised over FPT pairs per thread; with
// FPT = 4 the channel loop is bound by the FP32 pipe (24 FFMA2 issue slots per 7 weight loads were not).
constexpr int FPT = 4;
constexpr int FWD_THREADS = 128;

template <int HEADS, bool DENSE>
__global__ void __launch_bounds__(FWD_THREADS) relbias_fwd_kernel(RelArgs a) {
  extern __shared__ __align__(16) float sw[];
  pdl_wait(); pdl_launch();
  stage_weights<HEADS, DENSE>(sw, a, threadIdx.x, FWD_THREADS);
  __syncthreads();
  const u64* Wy2 = reinterpret_cast<const u64*>(sw);
  const u64* by2 = Wy2 + R2 * 4;
  const u64* Wr2 = by2 + R2;
  const float* br = reinterpret_cast<const float*>(Wr2 + R2 * HEADS);
  const unsigned p0 = (unsigned)FPT * (blockIdx.x * (unsigned)FWD_THREADS + threadIdx.x);
  if (p0 >= a.pairs) return;
  unsigned pr[FPT];                     // pairs past the end re-read pair p0 and are not stored
#pragma unroll
  for (int q = 0; q < FPT; ++q) pr[q] = p0 + q < a.pairs ? p0 + q : p0;
  u64 r[FPT][HEADS];
#pragma unroll
  for (int q = 0; q < FPT; ++q)
#pragma unroll
    for (int h = 0; h < HEADS; ++h) r[q][h] = pk2(br[h], 0.f);
  if (DENSE) {
#pragma unroll 2
    for (int c4 = 0; c4 < R / 4; ++c4) {
      u64 ea[FPT], eb[FPT];
#pragma unroll
      for (int q = 0; q < FPT; ++q) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(a.rel + (size_t)pr[q] * R) + c4);
        ea[q] = pk2(x.x, x.y); eb[q] = pk2(x.z, x.w);
      }
#pragma unroll
      for (int h = 0; h < HEADS; ++h) {
        const u64 w0 = Wr2[(2 * c4) * HEADS + h], w1 = Wr2[(2 * c4 + 1) * HEADS + h];
#pragma unroll
        for (int q = 0; q < FPT; ++q) r[q][h] = ffma2(w1, eb[q], ffma2(w0, ea[q], r[q][h]));
      }
    }
  } else {
    u64 gk[FPT][4];
#pragma unroll
    for (int q = 0; q < FPT; ++q) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(a.g4) + pr[q]);
      gk[q][0] = pk2(g.x, g.x); gk[q][1] = pk2(g.y, g.y); gk[q][2] = pk2(g.z, g.z); gk[q][3] = pk2(g.w, g.w);
    }
#pragma unroll 2
    for (int c2 = 0; c2 < R2; ++c2) {
      u64 e[FPT];
#pragma unroll
      for (int q = 0; q < FPT; ++q) e[q] = geo_channel_pair<HEADS>(Wy2, by2, gk[q], c2);
#pragma unroll
      for (int h = 0; h < HEADS; ++h) {
        const u64 w = Wr2[c2 * HEADS + h];
#pragma unroll
        for (int q = 0; q < FPT; ++q) r[q][h] = ffma2(w, e[q], r[q][h]);
      }
    }
  }
  const unsigned b0 = p0 / a.nn, ij0 = p0 - b0 * a.nn;
  if (p0 + FPT <= a.pairs && ij0 + FPT <= a.nn && (a.nn & 3u) == 0u) {     // all in one image, 16-byte aligned
    float* o = a.bias + ((size_t)b0 * HEADS) * a.nn + ij0;
#pragma unroll
    for (int h = 0; h < HEADS; ++h) {
      float v[FPT];
#pragma unroll
      for (int q = 0; q < FPT; ++q) v[q] = logf(fmaxf(fmaxf(sum2(r[q][h]), 0.f), 1e-6f));
      *reinterpret_cast<float4*>(o + (size_t)h * a.nn) = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
#pragma unroll
    for (int q = 0; q < FPT; ++q) {
      const unsigned p = p0 + q;
      if (p >= a.pairs) break;
      const unsigned b = p / a.nn, ij = p - b * a.nn;
      float* o = a.bias + ((size_t)b * HEADS) * a.nn + ij;
#pragma unroll
      for (int h = 0; h < HEADS; ++h) o[(size_t)h * a.nn] = logf(fmaxf(fmaxf(sum2(r[q][h]), 0.f), 1e-6f));
    }
  }
}

template <int HEADS, bool DENSE>
__global__ void __launch_bounds__(TILE, 4) relbias_bwd_kernel(RelArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* E = sm;                          // [R][EP]      e, transposed
  float* DpT = E + R * EP;                // [HEADS][EP]  d pre_r, transposed
  float* GT = DpT + HEADS * EP;           // [4][EP]      geometry, transposed
  float* sw = GT + 4 * EP;
  pdl_wait(); pdl_launch();
  const int t = threadIdx.x;
  stage_weights<HEADS, DENSE>(sw, a, t, TILE);
  const u64* Wy2 = reinterpret_cast<const u64*>(sw);
  const u64* by2 = Wy2 + R2 * 4;
  const u64* Wr2 = by2 + R2;
  const float* br = reinterpret_cast<const float*>(Wr2 + R2 * HEADS);
  const int c_own = t & 63, half = t >> 6;      // phase B: channel and which half of the tile's pairs
  u64 wcol[HEADS];                              // W_r[:, c_own], duplicated
#pragma unroll
  for (int h = 0; h < HEADS; ++h) {
    const float w = __ldg(a.Wr + h * R + c_own);
    wcol[h] = pk2(w, w);
  }
  const u64 one2 = pk2(1.f, 1.f);
  float accWr[HEADS], accWy[4] = {0.f, 0.f, 0.f, 0.f}, accby = 0.f, accbr = 0.f;
#pragma unroll
  for (int h = 0; h < HEADS; ++h) accWr[h] = 0.f;
  const unsigned ntiles = (a.pairs + TILE - 1) / TILE;
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();                            // weights staged / previous tile fully consumed
    // ---------------- phase A: thread = pair, packed over channel pairs
    const unsigned pair = tile * TILE + t;
    const bool live = pair < a.pairs;
    float db[HEADS];                            // issued first: their latency hides behind the channel loop
    {
      const unsigned b = live ? pair / a.nn : 0u, ij = live ? pair - b * a.nn : 0u;
      const float* dbp = a.dbias + ((size_t)b * HEADS) * a.nn + ij;
#pragma unroll
      for (int h = 0; h < HEADS; ++h) db[h] = live ? __ldg(dbp + (size_t)h * a.nn) : 0.f;
    }
    u64 r2[HEADS];
#pragma unroll
    for (int h = 0; h < HEADS; ++h) r2[h] = pk2(br[h], 0.f);
    if (DENSE) {
      const float4* e4 = reinterpret_cast<const float4*>(a.rel + (size_t)(live ? pair : 0u) * R);
#pragma unroll 4
      for (int c4 = 0; c4 < R / 4; ++c4) {
        float4 e = __ldg(e4 + c4);
        if (!live) e = make_float4(0.f, 0.f, 0.f, 0.f);
        E[(4 * c4 + 0) * EP + t] = e.x; E[(4 * c4 + 1) * EP + t] = e.y;
        E[(4 * c4 + 2) * EP + t] = e.z; E[(4 * c4 + 3) * EP + t] = e.w;
        const u64 ea = pk2(e.x, e.y), eb = pk2(e.z, e.w);
#pragma unroll
        for (int h = 0; h < HEADS; ++h)
          r2[h] = ffma2(Wr2[(2 * c4 + 1) * HEADS + h], eb, ffma2(Wr2[(2 * c4) * HEADS + h], ea, r2[h]));
      }
    } else {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) g = __ldg(reinterpret_cast<const float4*>(a.g4) + pair);
      GT[0 * EP + t] = g.x; GT[1 * EP + t] = g.y; GT[2 * EP + t] = g.z; GT[3 * EP + t] = g.w;
      const u64 gk[4] = {pk2(g.x, g.x), pk2(g.y, g.y), pk2(g.z, g.z), pk2(g.w, g.w)};
#pragma unroll 4
      for (int c2 = 0; c2 < R2; ++c2) {
        u64 e2 = geo_channel_pair<HEADS>(Wy2, by2, gk, c2);
        float lo, hi;
        unpk2(e2, lo, hi);
        if (!live) { lo = hi = 0.f; e2 = 0ull; }
        E[(2 * c2) * EP + t] = lo;
        E[(2 * c2 + 1) * EP + t] = hi;
#pragma unroll
        for (int h = 0; h < HEADS; ++h) r2[h] = ffma2(Wr2[c2 * HEADS + h], e2, r2[h]);
      }
    }
#pragma unroll
    for (int h = 0; h < HEADS; ++h) {
      const float r = sum2(r2[h]);
      DpT[h * EP + t] = (live && r > 1e-6f) ? __fdividef(db[h], r) : 0.f;     // relu and clamp both pass
    }
    __syncthreads();
    // ---------------- phase B: thread = channel c_own over its half of the pairs, packed over pair pairs
    u64 pWr[HEADS], pWy[4] = {0ull, 0ull, 0ull, 0ull}, pby = 0ull;
    float pbr = 0.f;
#pragma unroll
    for (int h = 0; h < HEADS; ++h) pWr[h] = 0ull;
    const int p0 = half * (TILE / 2);
#pragma unroll 2
    for (int p = p0; p < p0 + TILE / 2; p += 4) {
      const float4 ev = *reinterpret_cast<const float4*>(E + c_own * EP + p);
      const u64 ea = pk2(ev.x, ev.y), eb = pk2(ev.z, ev.w);
      u64 da = 0ull, dbq = 0ull;
#pragma unroll
      for (int h = 0; h < HEADS; ++h) {
        const float4 dv = *reinterpret_cast<const float4*>(DpT + h * EP + p);     // same address in every lane
        const u64 dpa = pk2(dv.x, dv.y), dpb = pk2(dv.z, dv.w);
        da = ffma2(wcol[h], dpa, da);
        dbq = ffma2(wcol[h], dpb, dbq);
        pWr[h] = ffma2(dpb, eb, ffma2(dpa, ea, pWr[h]));
      }
      if (DENSE) {            // d rel_embed[pair][c]: lanes = 32 consecutive channels of one pair
        float d0, d1, d2, d3;
        unpk2(da, d0, d1); unpk2(dbq, d2, d3);
        const size_t base = (size_t)tile * TILE + p;
        if (base + 0 < a.pairs) a.drel[(base + 0) * R + c_own] = d0;
        if (base + 1 < a.pairs) a.drel[(base + 1) * R + c_own] = d1;
        if (base + 2 < a.pairs) a.drel[(base + 2) * R + c_own] = d2;
        if (base + 3 < a.pairs) a.drel[(base + 3) * R + c_own] = d3;
      } else {
        float d0, d1, d2, d3;
        unpk2(da, d0, d1); unpk2(dbq, d2, d3);
        da = pk2(ev.x > 0.f ? d0 : 0.f, ev.y > 0.f ? d1 : 0.f);                   // through the relu of linear_y_rel
        dbq = pk2(ev.z > 0.f ? d2 : 0.f, ev.w > 0.f ? d3 : 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 gv = *reinterpret_cast<const float4*>(GT + k * EP + p);
          pWy[k] = ffma2(dbq, pk2(gv.z, gv.w), ffma2(da, pk2(gv.x, gv.y), pWy[k]));
        }
        pby = ffma2(dbq, one2, ffma2(da, one2, pby));
      }
      if (c_own < HEADS) {
        const float4 dv = *reinterpret_cast<const float4*>(DpT + c_own * EP + p);
        pbr += (dv.x + dv.y) + (dv.z + dv.w);
      }
    }
#pragma unroll
    for (int h = 0; h < HEADS; ++h) accWr[h] += sum2(pWr[h]);
#pragma unroll
    for (int k = 0; k < 4; ++k) accWy[k] += sum2(pWy[k]);
    accby += sum2(pby); accbr += pbr;
  }
#pragma unroll
  for (int h = 0; h < HEADS; ++h) atomicAdd(&a.dWr[h * R + c_own], accWr[h]);
  if (!DENSE) {
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(&a.dWy[c_own * 4 + k], accWy[k]);
    atomicAdd(&a.dby[c_own], accby);
  }
  if (c_own < HEADS) atomicAdd(&a.dbr[c_own], accbr);
}

template <int HEADS>
constexpr size_t bwd_smem_bytes() {
  return sizeof(float) * (R * EP + HEADS * EP + 4 * EP + sw_floats<HEADS>());
}

int check(int B, int N, int heads, int Rin, const float* rel, const float* g4, const float* Wy, const float* by,
          const float* Wr, const float* br) {
  MMNAS_CHECK_ARG(B >= 0 && N > 0, "relbias: bad sizes");
  MMNAS_CHECK_ARG(heads == 1 || heads == 2 || heads == 4 || heads == 8 || heads == 16, "relbias: heads must be 1, 2, 4, 8 or 16");
  MMNAS_CHECK_ARG(Rin == R, "relbias: REL_SIZE must be 64");
  MMNAS_CHECK_ARG((rel != nullptr) != (g4 != nullptr), "relbias: give exactly one of rel_embed / geometry");
  MMNAS_CHECK_ARG(!g4 || (Wy && by), "relbias: geometry mode needs linear_y_rel weights");
  MMNAS_CHECK_ARG(Wr && br, "relbias: linear_r weights missing");
  MMNAS_CHECK_ARG((double)B * N * N < 4.0e9, "relbias: too many pairs");
  return MMNAS_OK;
}

template <int HEADS>
int launch_fwd(const RelArgs& a, cudaStream_t s) {
  constexpr unsigned per_cta = FPT * FWD_THREADS;
  const unsigned grid = (a.pairs + per_cta - 1u) / per_cta;
  constexpr size_t smem = sizeof(float) * sw_floats<HEADS>();
  if (a.rel) MMNAS_CUDA(mmnas_launch(relbias_fwd_kernel<HEADS, true>, dim3(grid), dim3(FWD_THREADS), smem, s, a));
  else MMNAS_CUDA(mmnas_launch(relbias_fwd_kernel<HEADS, false>, dim3(grid), dim3(FWD_THREADS), smem, s, a));
  return MMNAS_OK;
}

template <int HEADS>
int launch_bwd(const RelArgs& a, cudaStream_t s) {
  constexpr size_t smem = bwd_smem_bytes<HEADS>();
  static bool attr_done = false;
  if (!attr_done) {
    MMNAS_CUDA(cudaFuncSetAttribute(relbias_bwd_kernel<HEADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMNAS_CUDA(cudaFuncSetAttribute(relbias_bwd_kernel<HEADS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const unsigned ntiles = (a.pairs + TILE - 1) / TILE;
  const unsigned grid = ntiles < 148u * 4u ? ntiles : 148u * 4u;
  if (a.rel) MMNAS_CUDA(mmnas_launch(relbias_bwd_kernel<HEADS, true>, dim3(grid), dim3(TILE), smem, s, a));
  else MMNAS_CUDA(mmnas_launch(relbias_bwd_kernel<HEADS, false>, dim3(grid), dim3(TILE), smem, s, a));
  return MMNAS_OK;
}

#define DISPATCH_HEADS(FN, heads, ...)            \
  switch (heads) {                                \
    case 1: return FN<1>(__VA_ARGS__);            \
    case 2: return FN<2>(__VA_ARGS__);            \
    case 4: return FN<4>(__VA_ARGS__);            \
    case 8: return FN<8>(__VA_ARGS__);            \
    default: return FN<16>(__VA_ARGS__);          \
  }

}  // namespace

int mmnas_relbias_fwd_mma(int B, int N, int heads, const float* g4, const float* Wy, const float* by, const float* Wr,
                          const float* br, float* bias, cudaStream_t s);
int mmnas_relbias_bwd_mma(int B, int N, int heads, const float* g4, const float* Wy, const float* by, const float* Wr,
                          const float* br, const float* dbias, float* dWy, float* dby, float* dWr, float* dbr,
                          cudaStream_t s);

extern "C" int mmnas_relbias_fwd(int mode, int B, int N, int heads, int Rin, const float* rel, const float* g4,
                                 const float* Wy, const float* by, const float* Wr, const float* br, float* bias,
                                 mmnas_stream stream) {
  int rc = check(B, N, heads, Rin, rel, g4, Wy, by, Wr, br);
  if (rc) return rc;
  MMNAS_CHECK_ARG(mode == 0 || mode == 1, "relbias_fwd: mode must be 0 (fp32 arithmetic) or 1 (tensor-core arithmetic)");
  if (B == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(bias, "relbias_fwd: null output");
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 1 && g4) {
    rc = mmnas_relbias_fwd_mma(B, N, heads, g4, Wy, by, Wr, br, bias, s);
    if (rc != MMNAS_ERR_UNSUPPORTED) return rc;
  }
  RelArgs a = {};
  a.B = B; a.N = N; a.heads = heads; a.nn = (unsigned)N * N; a.pairs = (unsigned)B * a.nn;
  a.rel = rel; a.g4 = g4; a.Wy = Wy; a.by = by; a.Wr = Wr; a.br = br; a.bias = bias;
  DISPATCH_HEADS(launch_fwd, heads, a, s)
}

extern "C" int mmnas_relbias_bwd(int mode, int B, int N, int heads, int Rin, const float* rel, const float* g4,
                                 const float* Wy, const float* by, const float* Wr, const float* br, const float* dbias,
                                 float* drel, float* dWy, float* dby, float* dWr, float* dbr, mmnas_stream stream) {
  int rc = check(B, N, heads, Rin, rel, g4, Wy, by, Wr, br);
  if (rc) return rc;
  MMNAS_CHECK_ARG(mode == 0 || mode == 1, "relbias_bwd: mode must be 0 (fp32 arithmetic) or 1 (tensor-core arithmetic)");
  if (B == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(dbias && dWr && dbr, "relbias_bwd: null buffer");
  MMNAS_CHECK_ARG(!rel || drel, "relbias_bwd: dense mode needs d rel_embed output");
  MMNAS_CHECK_ARG(!g4 || (dWy && dby), "relbias_bwd: geometry mode needs dWy/dby outputs");
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 1 && g4) {
    rc = mmnas_relbias_bwd_mma(B, N, heads, g4, Wy, by, Wr, br, dbias, dWy, dby, dWr, dbr, s);
    if (rc != MMNAS_ERR_UNSUPPORTED) return rc;
  }
  RelArgs a = {};
  a.B = B; a.N = N; a.heads = heads; a.nn = (unsigned)N * N; a.pairs = (unsigned)B * a.nn;
  a.rel = rel; a.g4 = g4; a.Wy = Wy; a.by = by; a.Wr = Wr; a.br = br; a.dbias = dbias; a.drel = drel; a.dWy = dWy; a.dby = dby; a.dWr = dWr; a.dbr = dbr;
  DISPATCH_HEADS(launch_bwd, heads, a, s)
}
