// Block tail: z = x + dropout(branch);  out = a_2 * (z - mean) / (std_unbiased + eps) + b_2
// (reference: modules.py:261-271 for the residual/dropout tail, :52-56 for the LayerNorm with
//  UNBIASED std and eps added to sigma).  HBM-bound: one warp per row, float4 accesses, the row
//  is read once from HBM (second and third passes hit L1), z overwrites the branch buffer so the
//  backward needs no extra activation.  fp32 statistics in both precision modes.
#include "common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int ROWS_PER_CTA = 8;   // 8 warps

__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
ln_fwd_kernel(int rows, int H, const float* __restrict__ x, float* __restrict__ branch,
              const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
              float* __restrict__ out, __nv_bfloat16* __restrict__ out16,
              float* __restrict__ mean_out, float* __restrict__ sigma_out, DropCfg drop) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROWS_PER_CTA + warp;
  if (row >= rows) return;
  const long base = (long)row * H;
  const bool use_drop = drop.state != nullptr && drop.thresh > 0;
  uint64_t key = use_drop ? drop_key(drop) : 0;
  const int nvec = H >> 2;
  // pass 1: z = x + drop(branch), written back over branch; accumulate the row sum
  float s = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float4 b = *reinterpret_cast<const float4*>(branch + base + 4 * v);
    if (use_drop) {
      uint64_t r = mmnas_mix64(key ^ ((uint64_t)((base >> 2) + v) * 0x9E3779B97F4A7C15ull));
      b.x *= ((unsigned)(r) & 0xFFFFu) < drop.thresh ? 0.f : drop.scale;
      b.y *= ((unsigned)(r >> 16) & 0xFFFFu) < drop.thresh ? 0.f : drop.scale;
      b.z *= ((unsigned)(r >> 32) & 0xFFFFu) < drop.thresh ? 0.f : drop.scale;
      b.w *= ((unsigned)(r >> 48) & 0xFFFFu) < drop.thresh ? 0.f : drop.scale;
    }
    if (x) {
      float4 xv = *reinterpret_cast<const float4*>(x + base + 4 * v);
      b.x += xv.x; b.y += xv.y; b.z += xv.z; b.w += xv.w;
    }
    *reinterpret_cast<float4*>(branch + base + 4 * v) = b;
    s += (b.x + b.y) + (b.z + b.w);
  }
  if (!gamma) {   // norm disabled: out = z
    __syncwarp();
    for (int v = lane; v < nvec; v += 32) {
      float4 z = *reinterpret_cast<const float4*>(branch + base + 4 * v);
      *reinterpret_cast<float4*>(out + base + 4 * v) = z;
      if (out16) {
        __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(out16 + base + 4 * v);
        o[0] = __floats2bfloat162_rn(z.x, z.y);
        o[1] = __floats2bfloat162_rn(z.z, z.w);
      }
    }
    return;
  }
  const float mean = warp_sum(s) / (float)H;
  __syncwarp();
  // pass 2: centred sum of squares (two-pass, matches torch.std's numerics)
  float q = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float4 z = *reinterpret_cast<const float4*>(branch + base + 4 * v);
    float a = z.x - mean, b = z.y - mean, c = z.z - mean, d = z.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float sigma = sqrtf(warp_sum(q) / (float)(H - 1));
  const float t = 1.f / (sigma + eps);
  if (lane == 0) { mean_out[row] = mean; sigma_out[row] = sigma; }
  // pass 3: normalise
  for (int v = lane; v < nvec; v += 32) {
    float4 z = *reinterpret_cast<const float4*>(branch + base + 4 * v);
    float4 g = *reinterpret_cast<const float4*>(gamma + 4 * v);
    float4 bb = *reinterpret_cast<const float4*>(beta + 4 * v);
    float4 o;
    o.x = g.x * (z.x - mean) * t + bb.x;
    o.y = g.y * (z.y - mean) * t + bb.y;
    o.z = g.z * (z.z - mean) * t + bb.z;
    o.w = g.w * (z.w - mean) * t + bb.w;
    *reinterpret_cast<float4*>(out + base + 4 * v) = o;
    if (out16) {
      __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(out16 + base + 4 * v);
      p[0] = __floats2bfloat162_rn(o.x, o.y);
      p[1] = __floats2bfloat162_rn(o.z, o.w);
    }
  }
}

// Backward.  With c = z - mean, t = 1/(sigma+eps), g = dout * a_2:
//   dL/dc_i = g_i t - t^2 (sum_j g_j c_j) c_i / ((H-1) sigma);   dz_i = dL/dc_i - t * mean(g)
// da_2 += dout * c * t (per column), db_2 += dout.  dbranch = dz * dropout multiplier.
constexpr int MAXV = 8;   // float4 groups per lane: supports H <= 1024

template <typename TB>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
ln_bwd_kernel(int rows, int H, const float* __restrict__ dout, const float* __restrict__ z,
              const float* __restrict__ mean_in, const float* __restrict__ sigma_in,
              const float* __restrict__ gamma, float eps, float* __restrict__ dz, TB* __restrict__ dbranch,
              float* __restrict__ dgamma, float* __restrict__ dbeta, DropCfg drop) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool norm = gamma != nullptr;
  const bool use_drop = drop.state != nullptr && drop.thresh > 0;
  const uint64_t key = use_drop ? drop_key(drop) : 0;
  const int nvec = H >> 2;
  float4 ag[MAXV], ab[MAXV];   // this lane's column partials of dgamma / dbeta over all its rows
#pragma unroll
  for (int k = 0; k < MAXV; ++k) { ag[k] = make_float4(0.f, 0.f, 0.f, 0.f); ab[k] = ag[k]; }

  for (int row = blockIdx.x * ROWS_PER_CTA + warp; row < rows; row += gridDim.x * ROWS_PER_CTA) {
    const long base = (long)row * H;
    float mean = 0.f, sigma = 1.f, t = 1.f, sg = 0.f, sgc = 0.f;
    if (norm) {
      mean = mean_in[row]; sigma = sigma_in[row]; t = 1.f / (sigma + eps);
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        int v = lane + 32 * k;
        if (v < nvec) {
          float4 d = *reinterpret_cast<const float4*>(dout + base + 4 * v);
          float4 zz = *reinterpret_cast<const float4*>(z + base + 4 * v);
          float4 g = *reinterpret_cast<const float4*>(gamma + 4 * v);
          float g0 = d.x * g.x, g1 = d.y * g.y, g2 = d.z * g.z, g3 = d.w * g.w;
          sg += (g0 + g1) + (g2 + g3);
          sgc += (g0 * (zz.x - mean) + g1 * (zz.y - mean)) + (g2 * (zz.z - mean) + g3 * (zz.w - mean));
        }
      }
      sg = warp_sum(sg);
      sgc = warp_sum(sgc);
    }
    const float k1 = norm ? t * sg / (float)H : 0.f;
    const float k2 = (norm && sigma > 0.f) ? t * t * sgc / ((float)(H - 1) * sigma) : 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      int v = lane + 32 * k;
      if (v < nvec) {
        float4 d = *reinterpret_cast<const float4*>(dout + base + 4 * v);
        float4 r;
        if (norm) {
          float4 zz = *reinterpret_cast<const float4*>(z + base + 4 * v);
          float4 g = *reinterpret_cast<const float4*>(gamma + 4 * v);
          float c0 = zz.x - mean, c1 = zz.y - mean, c2 = zz.z - mean, c3 = zz.w - mean;
          r.x = d.x * g.x * t - k2 * c0 - k1;
          r.y = d.y * g.y * t - k2 * c1 - k1;
          r.z = d.z * g.z * t - k2 * c2 - k1;
          r.w = d.w * g.w * t - k2 * c3 - k1;
          ag[k].x += d.x * c0 * t; ag[k].y += d.y * c1 * t; ag[k].z += d.z * c2 * t; ag[k].w += d.w * c3 * t;
          ab[k].x += d.x; ab[k].y += d.y; ab[k].z += d.z; ab[k].w += d.w;
        } else {
          r = d;
        }
        if (dz) *reinterpret_cast<float4*>(dz + base + 4 * v) = r;
        if (dbranch) {
          if (use_drop) {
            uint64_t rr = mmnas_mix64(key ^ ((uint64_t)((base >> 2) + v) * 0x9E3779B97F4A7C15ull));
            r.x *= ((unsigned)(rr) & 0xFFFFu) < drop.thresh ? 0.f : drop.scale;
            r.y *= ((unsigned)(rr >> 16) & 0xFFFFu) < drop.thresh ? 0.f : drop.scale;
            r.z *= ((unsigned)(rr >> 32) & 0xFFFFu) < drop.thresh ? 0.f : drop.scale;
            r.w *= ((unsigned)(rr >> 48) & 0xFFFFu) < drop.thresh ? 0.f : drop.scale;
          }
          TB* p = dbranch + base + 4 * v;
          p[0] = from_f32<TB>(r.x); p[1] = from_f32<TB>(r.y); p[2] = from_f32<TB>(r.z); p[3] = from_f32<TB>(r.w);
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
    for (int k = 0; k < MAXV; ++k) {
      int v = lane + 32 * k;
      if (v < nvec) {
        atomicAdd(&red[4 * v + 0], ag[k].x); atomicAdd(&red[4 * v + 1], ag[k].y);
        atomicAdd(&red[4 * v + 2], ag[k].z); atomicAdd(&red[4 * v + 3], ag[k].w);
        atomicAdd(&red[H + 4 * v + 0], ab[k].x); atomicAdd(&red[H + 4 * v + 1], ab[k].y);
        atomicAdd(&red[H + 4 * v + 2], ab[k].z); atomicAdd(&red[H + 4 * v + 3], ab[k].w);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
      atomicAdd(&dgamma[i], red[i]);
      atomicAdd(&dbeta[i], red[H + i]);
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
  ln_fwd_kernel<<<ceil_div(rows, ROWS_PER_CTA), ROWS_PER_CTA * 32, 0, (cudaStream_t)stream>>>(
      rows, H, x, branch, gamma, beta, eps, out, (__nv_bfloat16*)out_bf16, mean, sigma, make_drop(rng_state, salt, p));
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
  if (ctas > 148 * 4) ctas = 148 * 4;   // grid-stride over rows: column partials stay in registers
  dim3 grid(ctas), block(ROWS_PER_CTA * 32);
  cudaStream_t s = (cudaStream_t)stream;
  if (dbranch_dtype == 0)
    ln_bwd_kernel<float><<<grid, block, smem, s>>>(rows, H, dout, z, mean, sigma, gamma, eps, dz, (float*)dbranch,
                                                   dgamma, dbeta, d);
  else
    ln_bwd_kernel<__nv_bfloat16><<<grid, block, smem, s>>>(rows, H, dout, z, mean, sigma, gamma, eps, dz,
                                                           (__nv_bfloat16*)dbranch, dgamma, dbeta, d);
  MMNAS_LAUNCH_CHECK();
  return MMNAS_OK;
}
