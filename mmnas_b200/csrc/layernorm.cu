// Block tail: z = x + dropout(branch);  out = a_2 * (z - mean) / (std_unbiased + eps) + b_2
// (reference: modules.py:261-271 for the residual/dropout tail, :52-56 for the LayerNorm with
//  UNBIASED std and eps added to sigma).  HBM-bound: one warp per row, float4 accesses, the row
//  is read once from HBM (second and third passes hit L1), z overwrites the branch buffer so the
//  backward needs no extra activation.  fp32 statistics in both precision modes.
#include "common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int ROWS_PER_CTA = 8;   // 8 warps
constexpr int MAXV = 8;           // float4 groups per lane: supports H <= 1024

__device__ __forceinline__ float4 drop4_mult(uint64_t key, uint64_t group, unsigned thresh, float scale) {
  const uint64_t r = mmnas_mix64(key ^ (group * 0x9E3779B97F4A7C15ull));
  float4 m;
  m.x = ((unsigned)(r) & 0xFFFFu) < thresh ? 0.f : scale;
  m.y = ((unsigned)(r >> 16) & 0xFFFFu) < thresh ? 0.f : scale;
  m.z = ((unsigned)(r >> 32) & 0xFFFFu) < thresh ? 0.f : scale;
  m.w = ((unsigned)(r >> 48) & 0xFFFFu) < thresh ? 0.f : scale;
  return m;
}

// NV = float4 groups per lane (H = 128 * NV exactly, or H <= 128 * NV with guards when !EXACT).  The row lives in
// registers: HBM is touched once per operand.
template <int NV, bool EXACT>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
ln_fwd_kernel(int rows, int H, const float* __restrict__ x, float* __restrict__ branch,
              const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
              float* __restrict__ out, __nv_bfloat16* __restrict__ out16,
              float* __restrict__ mean_out, float* __restrict__ sigma_out, DropCfg drop) {
  pdl_wait(); pdl_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROWS_PER_CTA + warp;
  if (row >= rows) return;
  const long base = (long)row * H;
  const bool use_drop = drop.state != nullptr && drop.thresh > 0;
  const uint64_t key = use_drop ? drop_key(drop) : 0;
  const int nvec = H >> 2;
  float4 z[NV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    z[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EXACT || v < nvec) {
      float4 b = *reinterpret_cast<const float4*>(branch + base + 4 * v);
      if (use_drop) {
        const float4 m = drop4_mult(key, (uint64_t)((base >> 2) + v), drop.thresh, drop.scale);
        b.x *= m.x; b.y *= m.y; b.z *= m.z; b.w *= m.w;
      }
      if (x) {
        const float4 xv = *reinterpret_cast<const float4*>(x + base + 4 * v);
        b.x += xv.x; b.y += xv.y; b.z += xv.z; b.w += xv.w;
      }
      z[k] = b;
      s += (b.x + b.y) + (b.z + b.w);
    }
  }
  float mean = 0.f, t = 1.f;
  if (gamma) {
    mean = warp_sum(s) / (float)H;
    float q = 0.f;          // centred sum of squares (two-pass, matches torch.std's numerics)
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (EXACT || lane + 32 * k < nvec) {
        const float a = z[k].x - mean, b = z[k].y - mean, c = z[k].z - mean, d = z[k].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    const float sigma = sqrtf(warp_sum(q) / (float)(H - 1));
    t = 1.f / (sigma + eps);
    if (lane == 0) { mean_out[row] = mean; sigma_out[row] = sigma; }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (EXACT || v < nvec) {
      *reinterpret_cast<float4*>(branch + base + 4 * v) = z[k];       // z saved for the backward
      float4 o = z[k];
      if (gamma) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + v);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(beta) + v);
        o.x = g.x * (z[k].x - mean) * t + bb.x;
        o.y = g.y * (z[k].y - mean) * t + bb.y;
        o.z = g.z * (z[k].z - mean) * t + bb.z;
        o.w = g.w * (z[k].w - mean) * t + bb.w;
      }
      *reinterpret_cast<float4*>(out + base + 4 * v) = o;
      if (out16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
        uint2 pk;
        pk.x = *reinterpret_cast<unsigned*>(&lo);
        pk.y = *reinterpret_cast<unsigned*>(&hi);
        *reinterpret_cast<uint2*>(out16 + base + 4 * v) = pk;
      }
    }
  }
}

// Backward.  With c = z - mean, t = 1/(sigma+eps), g = dout * a_2:
//   dL/dc_i = g_i t - t^2 (sum_j g_j c_j) c_i / ((H-1) sigma);   dz_i = dL/dc_i - t * mean(g)
// da_2 += dout * c * t (per column), db_2 += dout.  dbranch = dz * dropout multiplier.
template <typename TB, int NV, bool EXACT>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
ln_bwd_kernel(int rows, int H, const float* __restrict__ dout, const float* __restrict__ z,
              const float* __restrict__ mean_in, const float* __restrict__ sigma_in,
              const float* __restrict__ gamma, float eps, float* __restrict__ dz, TB* __restrict__ dbranch,
              float* __restrict__ dgamma, float* __restrict__ dbeta, DropCfg drop) {
  pdl_wait(); pdl_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool norm = gamma != nullptr;
  const bool use_drop = drop.state != nullptr && drop.thresh > 0;
  const uint64_t key = use_drop ? drop_key(drop) : 0;
  const int nvec = H >> 2;
  float4 ag[NV], ab[NV], gm[NV];   // this lane's column partials of dgamma / dbeta over all its rows; gamma
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    ag[k] = make_float4(0.f, 0.f, 0.f, 0.f); ab[k] = ag[k]; gm[k] = ag[k];
    if (norm && (EXACT || lane + 32 * k < nvec)) gm[k] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * k);
  }
  for (int row = blockIdx.x * ROWS_PER_CTA + warp; row < rows; row += gridDim.x * ROWS_PER_CTA) {
    const long base = (long)row * H;
    float4 d[NV], c[NV];
    float mean = 0.f, sigma = 1.f, t = 1.f, sg = 0.f, sgc = 0.f;
    if (norm) { mean = mean_in[row]; sigma = sigma_in[row]; t = 1.f / (sigma + eps); }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      d[k] = make_float4(0.f, 0.f, 0.f, 0.f); c[k] = d[k];
      if (EXACT || v < nvec) {
        d[k] = *reinterpret_cast<const float4*>(dout + base + 4 * v);
        if (norm) {
          const float4 zz = *reinterpret_cast<const float4*>(z + base + 4 * v);
          c[k] = make_float4(zz.x - mean, zz.y - mean, zz.z - mean, zz.w - mean);
          const float g0 = d[k].x * gm[k].x, g1 = d[k].y * gm[k].y, g2 = d[k].z * gm[k].z, g3 = d[k].w * gm[k].w;
          sg += (g0 + g1) + (g2 + g3);
          sgc += (g0 * c[k].x + g1 * c[k].y) + (g2 * c[k].z + g3 * c[k].w);
        }
      }
    }
    float k1 = 0.f, k2 = 0.f;
    if (norm) {
      sg = warp_sum(sg);
      sgc = warp_sum(sgc);
      k1 = t * sg / (float)H;
      k2 = sigma > 0.f ? t * t * sgc / ((float)(H - 1) * sigma) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (EXACT || v < nvec) {
        float4 r = d[k];
        if (norm) {
          r.x = d[k].x * gm[k].x * t - k2 * c[k].x - k1;
          r.y = d[k].y * gm[k].y * t - k2 * c[k].y - k1;
          r.z = d[k].z * gm[k].z * t - k2 * c[k].z - k1;
          r.w = d[k].w * gm[k].w * t - k2 * c[k].w - k1;
          ag[k].x += d[k].x * c[k].x * t; ag[k].y += d[k].y * c[k].y * t;
          ag[k].z += d[k].z * c[k].z * t; ag[k].w += d[k].w * c[k].w * t;
          ab[k].x += d[k].x; ab[k].y += d[k].y; ab[k].z += d[k].z; ab[k].w += d[k].w;
        }
        if (dz) *reinterpret_cast<float4*>(dz + base + 4 * v) = r;
        if (dbranch) {
          if (use_drop) {
            const float4 m = drop4_mult(key, (uint64_t)((base >> 2) + v), drop.thresh, drop.scale);
            r.x *= m.x; r.y *= m.y; r.z *= m.z; r.w *= m.w;
          }
          TB* p = dbranch + base + 4 * v;
          if (sizeof(TB) == 2) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned*>(&lo);
            pk.y = *reinterpret_cast<unsigned*>(&hi);
            *reinterpret_cast<uint2*>(p) = pk;
          } else {
            *reinterpret_cast<float4*>(p) = r;
          }
        }
      }
    }
  }
  if (norm) {
    // combine the 8 warps of the CTA in shared memory, then one global atomic per column per CTA
    extern __shared__ float red[];   // [2][H]
    for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (EXACT || v < nvec) {
        atomicAdd(&red[4 * v + 0], ag[k].x); atomicAdd(&red[4 * v + 1], ag[k].y);
        atomicAdd(&red[4 * v + 2], ag[k].z); atomicAdd(&red[4 * v + 3], ag[k].w);
        atomicAdd(&red[H + 4 * v + 0], ab[k].x); atomicAdd(&red[H + 4 * v + 1], ab[k].y);
        atomicAdd(&red[H + 4 * v + 2], ab[k].z); atomicAdd(&red[H + 4 * v + 3], ab[k].w);
      }
    }
    __syncthreads();
    // 296 CTAs x 2H same-address atomics serialise in L2: four columns per reduction op when the buffers allow it
    if ((((uintptr_t)dgamma | (uintptr_t)dbeta) & 15) == 0) {
      for (int i = threadIdx.x; i < (H >> 2); i += blockDim.x) {
        const float4 a = *reinterpret_cast<const float4*>(red + 4 * i), b = *reinterpret_cast<const float4*>(red + H + 4 * i);
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dgamma + 4 * i), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbeta + 4 * i), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
      }
    } else {
      for (int i = threadIdx.x; i < H; i += blockDim.x) {
        atomicAdd(&dgamma[i], red[i]);
        atomicAdd(&dbeta[i], red[H + i]);
      }
    }
  }
}

}  // namespace

static DropCfg make_drop(const unsigned long long* st, unsigned long long salt, float p) {
  DropCfg d;
  d.state = (p > 0.f) ? st : nullptr;
  d.salt = salt;
  d.thresh = (unsigned)(p * 65536.f + 0.5f);
  d.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  return d;
}

extern "C" int mmnas_ln_residual_fwd(int rows, int H, const float* x, float* branch, const float* gamma,
                                     const float* beta, float eps, float* out, void* out_bf16, float* mean,
                                     float* sigma, const unsigned long long* rng_state, unsigned long long salt,
                                     float p, mmnas_stream stream) {
  MMNAS_CHECK_ARG(rows >= 0 && H > 1 && (H % 4) == 0, "ln_residual_fwd: H must be a multiple of 4");
  if (rows == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(branch && out, "ln_residual_fwd: null buffer");
  MMNAS_CHECK_ARG((gamma == nullptr) == (beta == nullptr), "ln_residual_fwd: gamma/beta must both be given");
  MMNAS_CHECK_ARG(!gamma || (mean && sigma), "ln_residual_fwd: statistics buffers required with norm");
  MMNAS_CHECK_ARG(H <= 128 * MAXV, "ln_residual_fwd: H must be <= 1024");
  const dim3 grid(ceil_div(rows, ROWS_PER_CTA)), block(ROWS_PER_CTA * 32);
  cudaStream_t s = (cudaStream_t)stream;
  const DropCfg d = make_drop(rng_state, salt, p);
  __nv_bfloat16* o16 = (__nv_bfloat16*)out_bf16;
#define LN_FWD(NV, EX) MMNAS_CUDA(mmnas_launch(ln_fwd_kernel<NV, EX>, grid, block, 0, s, rows, H, x, branch, gamma, beta, eps, out, o16, mean, sigma, d))
  if (H == 256) LN_FWD(2, true);
  else if (H == 512) LN_FWD(4, true);
  else if (H == 1024) LN_FWD(8, true);
  else if (H <= 128) LN_FWD(1, false);
  else if (H <= 512) LN_FWD(4, false);
  else LN_FWD(8, false);
#undef LN_FWD
  MMNAS_LAUNCH_CHECK();
  return MMNAS_OK;
}

extern "C" int mmnas_ln_residual_bwd(int rows, int H, const float* dout, const float* z, const float* mean,
                                     const float* sigma, const float* gamma, float eps, float* dz, void* dbranch,
                                     int dbranch_dtype, float* dgamma, float* dbeta,
                                     const unsigned long long* rng_state, unsigned long long salt, float p,
                                     mmnas_stream stream) {
  MMNAS_CHECK_ARG(rows >= 0 && H > 1 && (H % 4) == 0 && H <= 128 * MAXV,
                  "ln_residual_bwd: H must be a multiple of 4 and <= 1024");
  if (rows == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(dout, "ln_residual_bwd: null dout");
  MMNAS_CHECK_ARG(!gamma || (z && mean && sigma && dgamma && dbeta), "ln_residual_bwd: norm buffers missing");
  MMNAS_CHECK_ARG(dbranch_dtype == 0 || dbranch_dtype == 1, "ln_residual_bwd: dtype");
  DropCfg d = make_drop(rng_state, salt, p);
  size_t smem = gamma ? 2 * (size_t)H * sizeof(float) : 0;
  int ctas = ceil_div(rows, ROWS_PER_CTA);
  if (ctas > 148 * 2) ctas = 148 * 2;   // one resident wave (2 CTAs/SM at ~120 registers); rows are grid-strided
  dim3 grid(ctas), block(ROWS_PER_CTA * 32);
  cudaStream_t s = (cudaStream_t)stream;
#define LN_BWD(TB, NV, EX) MMNAS_CUDA(mmnas_launch(ln_bwd_kernel<TB, NV, EX>, grid, block, smem, s, rows, H, dout, z, mean, sigma, gamma, eps, dz, (TB*)dbranch, dgamma, dbeta, d))
#define LN_BWD_H(TB)                                  \
  do {                                                \
    if (H == 256) LN_BWD(TB, 2, true);                \
    else if (H == 512) LN_BWD(TB, 4, true);           \
    else if (H == 1024) LN_BWD(TB, 8, true);          \
    else if (H <= 128) LN_BWD(TB, 1, false);          \
    else if (H <= 512) LN_BWD(TB, 4, false);          \
    else LN_BWD(TB, 8, false);                        \
  } while (0)
  if (dbranch_dtype == 0) LN_BWD_H(float);
  else LN_BWD_H(__nv_bfloat16);
#undef LN_BWD_H
#undef LN_BWD
  MMNAS_LAUNCH_CHECK();
  return MMNAS_OK;
}
