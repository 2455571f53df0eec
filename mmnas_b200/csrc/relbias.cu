// RSA geometry path -> attention-logit bias.
//   bias[b,h,i,j] = log(max(relu(W_r . e_ij + b_r)[h], 1e-6))            (modules.py:231,235)
//   e_ij = rel_embed[b,i,j,:]                                  (dense mode: the reference's tensor)
//        = relu(W_y . g_ij + b_y),  g_ij = 4-d box log-geometry (geometry mode: full_vqa.py:82,103
//          folded in, so the [B,N,N,REL_SIZE] tensor never exists in HBM)
// Backward accumulates dW_r, db_r (and dW_y, db_y in geometry mode; d rel_embed in dense mode) from
// dbias = dS of the attention backward.  Thread per (i,j) pair for the pointwise part; the weight
// gradients are small GEMM-shaped reductions over the pairs, done per 256-pair tile in shared
// memory with per-thread register accumulators across a persistent grid (one global atomic per
// output per CTA).
#include "common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int MAXH = 16;     // heads
constexpr int MAXR = 64;     // REL_SIZE
constexpr int TILE = 256;    // pairs per tile == threads per CTA

struct RelArgs {
  int B, N, heads, R;
  long pairs;
  const float* rel;     // dense: [pairs, R]
  const float* g4;      // geometry: [pairs, 4]
  const float *Wy, *by; // [R,4], [R]
  const float *Wr, *br; // [heads,R], [heads]
  float* bias;          // [B, heads, N, N]
  const float* dbias;
  float* drel;
  float *dWy, *dby, *dWr, *dbr;
};

__device__ __forceinline__ long bias_index(const RelArgs& a, long pair, int hh) {
  const long nn = (long)a.N * a.N;
  const long b = pair / nn, ij = pair % nn;
  return (b * a.heads + hh) * nn + ij;
}

template <bool DENSE>
__global__ void __launch_bounds__(TILE) relbias_fwd_kernel(RelArgs a) {
  __shared__ float sWr[MAXH * MAXR], sbr[MAXH], sWy[MAXR * 4], sby[MAXR];
  for (int i = threadIdx.x; i < a.heads * a.R; i += TILE) sWr[i] = a.Wr[i];
  for (int i = threadIdx.x; i < a.heads; i += TILE) sbr[i] = a.br[i];
  if (!DENSE) {
    for (int i = threadIdx.x; i < a.R * 4; i += TILE) sWy[i] = a.Wy[i];
    for (int i = threadIdx.x; i < a.R; i += TILE) sby[i] = a.by[i];
  }
  __syncthreads();
  const long pair = (long)blockIdx.x * TILE + threadIdx.x;
  if (pair >= a.pairs) return;
  float r[MAXH];
#pragma unroll
  for (int hh = 0; hh < MAXH; ++hh) r[hh] = 0.f;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!DENSE) g = *reinterpret_cast<const float4*>(a.g4 + pair * 4);
  for (int c = 0; c < a.R; ++c) {
    float e;
    if (DENSE) e = a.rel[pair * a.R + c];
    else e = fmaxf(fmaf(sWy[c * 4 + 0], g.x, fmaf(sWy[c * 4 + 1], g.y, fmaf(sWy[c * 4 + 2], g.z, fmaf(sWy[c * 4 + 3], g.w, sby[c])))), 0.f);
#pragma unroll
    for (int hh = 0; hh < MAXH; ++hh)
      if (hh < a.heads) r[hh] = fmaf(sWr[hh * a.R + c], e, r[hh]);
  }
#pragma unroll
  for (int hh = 0; hh < MAXH; ++hh)
    if (hh < a.heads) a.bias[bias_index(a, pair, hh)] = logf(fmaxf(fmaxf(r[hh] + sbr[hh], 0.f), 1e-6f));
}

template <bool DENSE>
__global__ void __launch_bounds__(TILE) relbias_bwd_kernel(RelArgs a) {
  extern __shared__ float sm[];
  float* sWr = sm;                       // [heads][R]
  float* sbr = sWr + MAXH * MAXR;        // [heads]
  float* sWy = sbr + MAXH;               // [R][4]
  float* sby = sWy + MAXR * 4;           // [R]
  float* E = sby + MAXR;                 // [TILE][R]    e (post-ReLU in geometry mode)
  float* DE = E + TILE * MAXR;           // [TILE][R]    d pre_e (geometry) / d rel (dense)
  float* Dp = DE + TILE * MAXR;          // [TILE][MAXH] d pre_r
  float* G = Dp + TILE * MAXH;           // [TILE][4]
  const int t = threadIdx.x, R = a.R, heads = a.heads;
  for (int i = t; i < heads * R; i += TILE) sWr[i] = a.Wr[i];
  for (int i = t; i < heads; i += TILE) sbr[i] = a.br[i];
  if (!DENSE) {
    for (int i = t; i < R * 4; i += TILE) sWy[i] = a.Wy[i];
    for (int i = t; i < R; i += TILE) sby[i] = a.by[i];
  }
  // register accumulators: dWr[(t/64) + 4k][t%64] for k<4, dWy[t%64][t/64], dby[t] (t<R), dbr[t] (t<heads)
  float accWr[4] = {0.f, 0.f, 0.f, 0.f}, accWy = 0.f, accby = 0.f, accbr = 0.f;
  const int c_own = t % 64, q_own = t / 64;
  const long ntiles = (a.pairs + TILE - 1) / TILE;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    const long pair = tile * TILE + t;
    const bool live = pair < a.pairs;
    // ---- phase A: pointwise recompute + chain rule for this thread's pair
    if (DENSE) {   // coalesced tile load of rel
      const long base = tile * TILE * (long)R;
      const long lim = a.pairs * (long)R;
      for (int e = t; e < TILE * R; e += TILE) E[e] = (base + e < lim) ? a.rel[base + e] : 0.f;
      __syncthreads();
    }
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!DENSE && live) g = *reinterpret_cast<const float4*>(a.g4 + pair * 4);
    float r[MAXH];
#pragma unroll
    for (int hh = 0; hh < MAXH; ++hh) r[hh] = 0.f;
    // E is written/read by its owner thread with a rotated column order so lanes hit distinct banks
    for (int cc = 0; cc < R; ++cc) {
      const int c = DENSE ? cc : ((cc + t) & (R - 1));
      float e;
      if (DENSE) e = E[t * R + ((cc + t) & (R - 1))];
      else {
        e = fmaxf(fmaf(sWy[c * 4 + 0], g.x, fmaf(sWy[c * 4 + 1], g.y, fmaf(sWy[c * 4 + 2], g.z, fmaf(sWy[c * 4 + 3], g.w, sby[c])))), 0.f);
        E[t * R + c] = live ? e : 0.f;
      }
      const int cw = DENSE ? ((cc + t) & (R - 1)) : c;
#pragma unroll
      for (int hh = 0; hh < MAXH; ++hh)
        if (hh < heads) r[hh] = fmaf(sWr[hh * R + cw], e, r[hh]);
    }
    float dpre[MAXH];
#pragma unroll
    for (int hh = 0; hh < MAXH; ++hh) {
      dpre[hh] = 0.f;
      if (hh < heads && live) {
        const float rv = r[hh] + sbr[hh];
        if (rv > 1e-6f) dpre[hh] = a.dbias[bias_index(a, pair, hh)] / rv;   // clamp & relu both pass
      }
      if (hh < heads) Dp[t * MAXH + hh] = dpre[hh];
    }
    for (int cc = 0; cc < R; ++cc) {
      const int c = (cc + t) & (R - 1);
      float de = 0.f;
#pragma unroll
      for (int hh = 0; hh < MAXH; ++hh)
        if (hh < heads) de = fmaf(sWr[hh * R + c], dpre[hh], de);
      if (!DENSE) de = E[t * R + c] > 0.f ? de : 0.f;
      DE[t * R + c] = de;
    }
    if (!DENSE) *reinterpret_cast<float4*>(G + t * 4) = g;
    __syncthreads();
    if (DENSE) {   // coalesced store of d rel_embed
      const long base = tile * TILE * (long)R;
      const long lim = a.pairs * (long)R;
      for (int e = t; e < TILE * R; e += TILE)
        if (base + e < lim) a.drel[base + e] = DE[e];
    }
    // ---- phase B: reductions over the tile's pairs.  Two-level summation (tile-local partials, then the
    // running total) keeps fp32 round-off at the level of a blocked GEMM instead of a 70k-term serial sum.
    float pWr[4] = {0.f, 0.f, 0.f, 0.f}, pWy = 0.f, pby = 0.f, pbr = 0.f;
    for (int p = 0; p < TILE; ++p) {
      const float ev = E[p * R + c_own];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int hh = q_own + 4 * k;
        if (hh < heads) pWr[k] = fmaf(Dp[p * MAXH + hh], ev, pWr[k]);
      }
      if (!DENSE) {
        const float dev = DE[p * R + c_own];
        pWy = fmaf(dev, G[p * 4 + q_own], pWy);
        if (q_own == 0) pby += dev;
      }
      if (t < heads) pbr += Dp[p * MAXH + t];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) accWr[k] += pWr[k];
    accWy += pWy; accby += pby; accbr += pbr;
  }
  if (c_own < R) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int hh = q_own + 4 * k;
      if (hh < heads) atomicAdd(&a.dWr[hh * R + c_own], accWr[k]);
    }
    if (!DENSE) {
      atomicAdd(&a.dWy[c_own * 4 + q_own], accWy);
      if (q_own == 0) atomicAdd(&a.dby[c_own], accby);
    }
  }
  if (t < heads) atomicAdd(&a.dbr[t], accbr);
}

size_t bwd_smem_bytes() {
  return sizeof(float) * (MAXH * MAXR + MAXH + MAXR * 4 + MAXR + 2 * TILE * MAXR + TILE * MAXH + TILE * 4);
}

int check(int B, int N, int heads, int R, const float* rel, const float* g4, const float* Wy, const float* by,
          const float* Wr, const float* br) {
  MMNAS_CHECK_ARG(B >= 0 && N > 0, "relbias: bad sizes");
  MMNAS_CHECK_ARG(heads >= 1 && heads <= MAXH, "relbias: heads must be in [1,16]");
  MMNAS_CHECK_ARG(R == MAXR, "relbias: REL_SIZE must be 64");
  MMNAS_CHECK_ARG((rel != nullptr) != (g4 != nullptr), "relbias: give exactly one of rel_embed / geometry");
  MMNAS_CHECK_ARG(!g4 || (Wy && by), "relbias: geometry mode needs linear_y_rel weights");
  MMNAS_CHECK_ARG(Wr && br, "relbias: linear_r weights missing");
  return MMNAS_OK;
}

}  // namespace

extern "C" int mmnas_relbias_fwd(int B, int N, int heads, int R, const float* rel, const float* g4, const float* Wy,
                                 const float* by, const float* Wr, const float* br, float* bias,
                                 mmnas_stream stream) {
  int rc = check(B, N, heads, R, rel, g4, Wy, by, Wr, br);
  if (rc) return rc;
  if (B == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(bias, "relbias_fwd: null output");
  RelArgs a = {};
  a.B = B; a.N = N; a.heads = heads; a.R = R; a.pairs = (long)B * N * N;
  a.rel = rel; a.g4 = g4; a.Wy = Wy; a.by = by; a.Wr = Wr; a.br = br; a.bias = bias;
  const int grid = (int)((a.pairs + TILE - 1) / TILE);
  if (rel) relbias_fwd_kernel<true><<<grid, TILE, 0, (cudaStream_t)stream>>>(a);
  else relbias_fwd_kernel<false><<<grid, TILE, 0, (cudaStream_t)stream>>>(a);
  MMNAS_LAUNCH_CHECK();
  return MMNAS_OK;
}

extern "C" int mmnas_relbias_bwd(int B, int N, int heads, int R, const float* rel, const float* g4, const float* Wy,
                                 const float* by, const float* Wr, const float* br, const float* dbias, float* drel,
                                 float* dWy, float* dby, float* dWr, float* dbr, mmnas_stream stream) {
  int rc = check(B, N, heads, R, rel, g4, Wy, by, Wr, br);
  if (rc) return rc;
  if (B == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(dbias && dWr && dbr, "relbias_bwd: null buffer");
  MMNAS_CHECK_ARG(!rel || drel, "relbias_bwd: dense mode needs d rel_embed output");
  MMNAS_CHECK_ARG(!g4 || (dWy && dby), "relbias_bwd: geometry mode needs dWy/dby outputs");
  RelArgs a = {};
  a.B = B; a.N = N; a.heads = heads; a.R = R; a.pairs = (long)B * N * N;
  a.rel = rel; a.g4 = g4; a.Wy = Wy; a.by = by; a.Wr = Wr; a.br = br;
  a.dbias = dbias; a.drel = drel; a.dWy = dWy; a.dby = dby; a.dWr = dWr; a.dbr = dbr;
  const long ntiles = (a.pairs + TILE - 1) / TILE;
  const int grid = (int)(ntiles < 148 ? ntiles : 148);
  const size_t smem = bwd_smem_bytes();
  cudaStream_t s = (cudaStream_t)stream;
  static bool attr_done = false;
  if (!attr_done) {
    MMNAS_CUDA(cudaFuncSetAttribute(relbias_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMNAS_CUDA(cudaFuncSetAttribute(relbias_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  if (rel) relbias_bwd_kernel<true><<<grid, TILE, smem, s>>>(a);
  else relbias_bwd_kernel<false><<<grid, TILE, smem, s>>>(a);
  MMNAS_LAUNCH_CHECK();
  return MMNAS_OK;
}
