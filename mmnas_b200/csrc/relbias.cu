// RSA geometry path -> attention-logit bias.
//   bias[b,h,i,j] = log(max(relu(W_r . e_ij + b_r)[h], 1e-6))            (modules.py:231,235)
//   e_ij = rel_embed[b,i,j,:]                                  (dense mode: the reference's tensor)
//        = relu(W_y . g_ij + b_y),  g_ij = 4-d box log-geometry (geometry mode: full_vqa.py:82,103
//          folded in, so the [B,N,N,REL_SIZE] tensor never exists in HBM)
// Backward accumulates dW_r, db_r (and dW_y, db_y in geometry mode; d rel_embed in dense mode) from
// dbias = dS of the attention backward.
//
// Design (v2).  The layer weights (<= 5.5 KB) are staged into __constant__ memory per call, so the fully
// unrolled per-pair MLP uses FFMA with constant-bank operands: no weight loads at all.  Forward: one thread per
// (i,j) pair, no shared memory.  Backward: 128-pair tiles; phase A (thread = pair) recomputes e, r and the
// chain rule and parks e / d pre_e transposed in shared memory ([c][pair], padded so both the scalar writes of
// phase A and the 128-bit reads of phase B are bank-conflict free); phase B (thread = column c, two half-tiles)
// reduces the tile into 13 register accumulators per thread (dW_r[:,c], dW_y[c,:], db_y[c]) with a two-level
// sum; one global atomic per output per CTA at the end.  Persistent grid of 3 CTAs per SM.
// The constant staging makes these entry points single-stream per device (calls are stream-ordered).
#include "common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int MAXH = 16;     // heads
constexpr int R = 64;        // REL_SIZE
constexpr int TILE = 128;    // pairs per tile == threads per CTA
constexpr int EP = TILE + 4; // padded pair pitch of the transposed tiles

__constant__ float cWy[R * 4];
__constant__ float cby[R];
__constant__ float cWr[MAXH * R];
__constant__ float cbr[MAXH];

struct RelArgs {
  int B, N, heads;
  unsigned pairs, nn;
  const float* rel;     // dense: [pairs, R]
  const float* g4;      // geometry: [pairs, 4]
  float* bias;          // [B, heads, N, N]
  const float* dbias;
  float* drel;
  float *dWy, *dby, *dWr, *dbr;
};

template <int HEADS, bool DENSE>
__global__ void __launch_bounds__(256) relbias_fwd_kernel(RelArgs a) {
  pdl_wait(); pdl_launch();
  const unsigned pair = blockIdx.x * 256u + threadIdx.x;
  if (pair >= a.pairs) return;
  float r[HEADS];
#pragma unroll
  for (int h = 0; h < HEADS; ++h) r[h] = cbr[h];
  if (DENSE) {
    const float4* e4 = reinterpret_cast<const float4*>(a.rel + (size_t)pair * R);
#pragma unroll
    for (int c4 = 0; c4 < R / 4; ++c4) {
      const float4 e = __ldg(e4 + c4);
#pragma unroll
      for (int h = 0; h < HEADS; ++h) {
        r[h] = fmaf(cWr[h * R + 4 * c4 + 0], e.x, r[h]);
        r[h] = fmaf(cWr[h * R + 4 * c4 + 1], e.y, r[h]);
        r[h] = fmaf(cWr[h * R + 4 * c4 + 2], e.z, r[h]);
        r[h] = fmaf(cWr[h * R + 4 * c4 + 3], e.w, r[h]);
      }
    }
  } else {
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.g4) + pair);
#pragma unroll
    for (int c = 0; c < R; ++c) {
      const float e = fmaxf(fmaf(cWy[c * 4 + 0], g.x, fmaf(cWy[c * 4 + 1], g.y, fmaf(cWy[c * 4 + 2], g.z, fmaf(cWy[c * 4 + 3], g.w, cby[c])))), 0.f);
#pragma unroll
      for (int h = 0; h < HEADS; ++h) r[h] = fmaf(cWr[h * R + c], e, r[h]);
    }
  }
  const unsigned b = pair / a.nn, ij = pair - b * a.nn;
  float* out = a.bias + ((size_t)b * HEADS) * a.nn + ij;
#pragma unroll
  for (int h = 0; h < HEADS; ++h) out[(size_t)h * a.nn] = logf(fmaxf(fmaxf(r[h], 0.f), 1e-6f));
}

template <int HEADS, bool DENSE>
__global__ void __launch_bounds__(TILE) relbias_bwd_kernel(RelArgs a) {
  extern __shared__ float sm[];
  float* E = sm;                     // [R][EP]   e, transposed
  float* DE = E + R * EP;            // [R][EP]   d pre_e (geometry) / d rel (dense), transposed
  float* Dp = DE + R * EP;           // [TILE][HP] d pre_r
  constexpr int HP = HEADS < 4 ? 4 : HEADS;
  float* G = Dp + TILE * HP;         // [TILE][4]
  pdl_wait(); pdl_launch();
  const int t = threadIdx.x;
  const int c_own = t & 63, half = t >> 6;      // phase B: column and which half of the tile's pairs
  float accWr[HEADS], accWy[4] = {0.f, 0.f, 0.f, 0.f}, accby = 0.f, accbr = 0.f;
#pragma unroll
  for (int h = 0; h < HEADS; ++h) accWr[h] = 0.f;
  const unsigned ntiles = (a.pairs + TILE - 1) / TILE;
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    // ---------------- phase A: thread = pair
    const unsigned pair = tile * TILE + t;
    const bool live = pair < a.pairs;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    float r[HEADS];
#pragma unroll
    for (int h = 0; h < HEADS; ++h) r[h] = cbr[h];
    if (DENSE) {
      const float4* e4 = reinterpret_cast<const float4*>(a.rel + (size_t)pair * R);
#pragma unroll
      for (int c4 = 0; c4 < R / 4; ++c4) {
        const float4 e = live ? __ldg(e4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        E[(4 * c4 + 0) * EP + t] = e.x; E[(4 * c4 + 1) * EP + t] = e.y;
        E[(4 * c4 + 2) * EP + t] = e.z; E[(4 * c4 + 3) * EP + t] = e.w;
#pragma unroll
        for (int h = 0; h < HEADS; ++h) {
          r[h] = fmaf(cWr[h * R + 4 * c4 + 0], e.x, r[h]);
          r[h] = fmaf(cWr[h * R + 4 * c4 + 1], e.y, r[h]);
          r[h] = fmaf(cWr[h * R + 4 * c4 + 2], e.z, r[h]);
          r[h] = fmaf(cWr[h * R + 4 * c4 + 3], e.w, r[h]);
        }
      }
    } else {
      if (live) g = __ldg(reinterpret_cast<const float4*>(a.g4) + pair);
#pragma unroll
      for (int c = 0; c < R; ++c) {
        float e = fmaxf(fmaf(cWy[c * 4 + 0], g.x, fmaf(cWy[c * 4 + 1], g.y, fmaf(cWy[c * 4 + 2], g.z, fmaf(cWy[c * 4 + 3], g.w, cby[c])))), 0.f);
        e = live ? e : 0.f;
        E[c * EP + t] = e;
#pragma unroll
        for (int h = 0; h < HEADS; ++h) r[h] = fmaf(cWr[h * R + c], e, r[h]);
      }
      *reinterpret_cast<float4*>(G + t * 4) = g;
    }
    float dpre[HEADS];
    {
      const unsigned b = live ? pair / a.nn : 0u, ij = live ? pair - b * a.nn : 0u;
      const float* db = a.dbias + ((size_t)b * HEADS) * a.nn + ij;
#pragma unroll
      for (int h = 0; h < HEADS; ++h) {
        dpre[h] = (live && r[h] > 1e-6f) ? __ldg(db + (size_t)h * a.nn) / r[h] : 0.f;   // relu and clamp both pass
        Dp[t * HP + h] = dpre[h];
      }
    }
#pragma unroll
    for (int c = 0; c < R; ++c) {
      float de = 0.f;
#pragma unroll
      for (int h = 0; h < HEADS; ++h) de = fmaf(cWr[h * R + c], dpre[h], de);
      if (!DENSE) de = E[c * EP + t] > 0.f ? de : 0.f;
      DE[c * EP + t] = de;
    }
    __syncthreads();
    if (DENSE && live) {      // d rel_embed row of this pair
      float4* o4 = reinterpret_cast<float4*>(a.drel + (size_t)pair * R);
#pragma unroll
      for (int c4 = 0; c4 < R / 4; ++c4)
        o4[c4] = make_float4(DE[(4 * c4 + 0) * EP + t], DE[(4 * c4 + 1) * EP + t], DE[(4 * c4 + 2) * EP + t],
                             DE[(4 * c4 + 3) * EP + t]);
    }
    // ---------------- phase B: thread = column c_own over its half of the pairs; two-level summation
    float pWr[HEADS], pWy[4] = {0.f, 0.f, 0.f, 0.f}, pby = 0.f, pbr = 0.f;
#pragma unroll
    for (int h = 0; h < HEADS; ++h) pWr[h] = 0.f;
    const int p0 = half * (TILE / 2);
#pragma unroll 2
    for (int p = p0; p < p0 + TILE / 2; p += 4) {
      const float4 ev = *reinterpret_cast<const float4*>(E + c_own * EP + p);
      const float ee[4] = {ev.x, ev.y, ev.z, ev.w};
      float dd[4] = {0.f, 0.f, 0.f, 0.f};
      if (!DENSE) {
        const float4 dv = *reinterpret_cast<const float4*>(DE + c_own * EP + p);
        dd[0] = dv.x; dd[1] = dv.y; dd[2] = dv.z; dd[3] = dv.w;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int h = 0; h < HEADS; ++h) pWr[h] = fmaf(Dp[(p + u) * HP + h], ee[u], pWr[h]);
        if (!DENSE) {
          const float4 gv = *reinterpret_cast<const float4*>(G + (p + u) * 4);
          pWy[0] = fmaf(dd[u], gv.x, pWy[0]); pWy[1] = fmaf(dd[u], gv.y, pWy[1]);
          pWy[2] = fmaf(dd[u], gv.z, pWy[2]); pWy[3] = fmaf(dd[u], gv.w, pWy[3]);
          pby += dd[u];
        }
        if (c_own < HEADS) pbr += Dp[(p + u) * HP + c_own];
      }
    }
#pragma unroll
    for (int h = 0; h < HEADS; ++h) accWr[h] += pWr[h];
#pragma unroll
    for (int k = 0; k < 4; ++k) accWy[k] += pWy[k];
    accby += pby; accbr += pbr;
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
  return sizeof(float) * (2 * R * EP + TILE * (HEADS < 4 ? 4 : HEADS) + TILE * 4);
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

int stage_weights(int heads, const float* Wy, const float* by, const float* Wr, const float* br, cudaStream_t s) {
  if (Wy) {
    MMNAS_CUDA(cudaMemcpyToSymbolAsync(cWy, Wy, sizeof(float) * R * 4, 0, cudaMemcpyDeviceToDevice, s));
    MMNAS_CUDA(cudaMemcpyToSymbolAsync(cby, by, sizeof(float) * R, 0, cudaMemcpyDeviceToDevice, s));
  }
  MMNAS_CUDA(cudaMemcpyToSymbolAsync(cWr, Wr, sizeof(float) * heads * R, 0, cudaMemcpyDeviceToDevice, s));
  MMNAS_CUDA(cudaMemcpyToSymbolAsync(cbr, br, sizeof(float) * heads, 0, cudaMemcpyDeviceToDevice, s));
  return MMNAS_OK;
}

template <int HEADS>
int launch_fwd(const RelArgs& a, cudaStream_t s) {
  const unsigned grid = (a.pairs + 255u) / 256u;
  if (a.rel) MMNAS_CUDA(mmnas_launch(relbias_fwd_kernel<HEADS, true>, dim3(grid), dim3(256), 0, s, a));
  else MMNAS_CUDA(mmnas_launch(relbias_fwd_kernel<HEADS, false>, dim3(grid), dim3(256), 0, s, a));
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
  const unsigned grid = ntiles < 148u * 3u ? ntiles : 148u * 3u;
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

extern "C" int mmnas_relbias_fwd(int B, int N, int heads, int Rin, const float* rel, const float* g4, const float* Wy,
                                 const float* by, const float* Wr, const float* br, float* bias,
                                 mmnas_stream stream) {
  int rc = check(B, N, heads, Rin, rel, g4, Wy, by, Wr, br);
  if (rc) return rc;
  if (B == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(bias, "relbias_fwd: null output");
  cudaStream_t s = (cudaStream_t)stream;
  rc = stage_weights(heads, g4 ? Wy : nullptr, by, Wr, br, s);
  if (rc) return rc;
  RelArgs a = {};
  a.B = B; a.N = N; a.heads = heads; a.nn = (unsigned)N * N; a.pairs = (unsigned)B * a.nn;
  a.rel = rel; a.g4 = g4; a.bias = bias;
  DISPATCH_HEADS(launch_fwd, heads, a, s)
}

extern "C" int mmnas_relbias_bwd(int B, int N, int heads, int Rin, const float* rel, const float* g4, const float* Wy,
                                 const float* by, const float* Wr, const float* br, const float* dbias, float* drel,
                                 float* dWy, float* dby, float* dWr, float* dbr, mmnas_stream stream) {
  int rc = check(B, N, heads, Rin, rel, g4, Wy, by, Wr, br);
  if (rc) return rc;
  if (B == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(dbias && dWr && dbr, "relbias_bwd: null buffer");
  MMNAS_CHECK_ARG(!rel || drel, "relbias_bwd: dense mode needs d rel_embed output");
  MMNAS_CHECK_ARG(!g4 || (dWy && dby), "relbias_bwd: geometry mode needs dWy/dby outputs");
  cudaStream_t s = (cudaStream_t)stream;
  rc = stage_weights(heads, g4 ? Wy : nullptr, by, Wr, br, s);
  if (rc) return rc;
  RelArgs a = {};
  a.B = B; a.N = N; a.heads = heads; a.nn = (unsigned)N * N; a.pairs = (unsigned)B * a.nn;
  a.rel = rel; a.g4 = g4; a.dbias = dbias; a.drel = drel; a.dWy = dWy; a.dby = dby; a.dWr = dWr; a.dbr = dbr;
  DISPATCH_HEADS(launch_bwd, heads, a, s)
}
