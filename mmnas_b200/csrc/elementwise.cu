// HBM-bound helpers of the hot path: fp32->bf16 casts (operand shadows for the tensor-core GEMMs),
// column sums (bias gradients), the supernet mixed-op accumulate and its alpha-gate gradient
// (mixed.py:60-68 and the autograd rule SURVEY §8 a11), and the dropout step counter.
// All are vectorised (float4 / 8-byte bf16x4), grid-stride, grid = k x 148 SMs.
#include <atomic>
#include <cstdlib>
#include "common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int EW_THREADS = 256;
inline int ew_grid(long nvec) {
  long want = (nvec + EW_THREADS - 1) / EW_THREADS;
  long cap = 148 * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

__global__ void __launch_bounds__(EW_THREADS) cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long n) {
  pdl_wait(); pdl_launch();
  const long nvec = n >> 2;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
    float4 f = *reinterpret_cast<const float4*>(src + 4 * v);
    __nv_bfloat162 lo = __floats2bfloat162_rn(f.x, f.y), hi = __floats2bfloat162_rn(f.z, f.w);
    uint2 pk;
    pk.x = *reinterpret_cast<unsigned*>(&lo);
    pk.y = *reinterpret_cast<unsigned*>(&hi);
    *reinterpret_cast<uint2*>(dst + 4 * v) = pk;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    long i = (nvec << 2) + threadIdx.x;
    dst[i] = __float2bfloat16_rn(src[i]);
  }
}

// out[c] = sum_r x[r, c].  A thread owns 8 consecutive columns (one 16-byte bf16 load / two float4 loads per row),
// a CTA = 32 column groups x 8 row lanes over a slab of rows; smem combine; one atomic per column per CTA.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, int rows, int cols, long ld, float* __restrict__ out) {
  __shared__ float red[8][32 * 8 + 8];
  pdl_wait(); pdl_launch();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + tx) * 8;
  float s[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) s[u] = 0.f;
  const bool vec = (c0 + 8 <= cols) && (ld % 8 == 0);
  if (c0 < cols)
    for (int r = blockIdx.y * 8 + ty; r < rows; r += gridDim.y * 8) {
      const T* p = x + (long)r * ld + c0;
      if (vec) {
        if (sizeof(T) == 2) {
          const uint4 v = *reinterpret_cast<const uint4*>(p);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
          for (int u = 0; u < 4; ++u) { const float2 f = __bfloat1622float2(h[u]); s[2 * u] += f.x; s[2 * u + 1] += f.y; }
        } else {
          const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
          s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w; s[4] += b.x; s[5] += b.y; s[6] += b.z; s[7] += b.w;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (c0 + u < cols) s[u] += to_f32<T>(p[u]);
      }
    }
#pragma unroll
  for (int u = 0; u < 8; ++u) red[ty][tx * 8 + u] = s[u];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < cols) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += red[k][threadIdx.x];
    atomicAdd(&out[c], tot);
  }
}

constexpr int MAXK = 8;
struct MixedArgs {
  int K;
  const float* o[MAXK];
  float* d_o[MAXK];
  const float* gate;
  const float* dout;
  float* out;
  float* gate_grad;
  long n;
};

__global__ void __launch_bounds__(EW_THREADS) mixed_accum_kernel(MixedArgs a) {
  pdl_wait(); pdl_launch();
  float g[MAXK];
#pragma unroll
  for (int k = 0; k < MAXK; ++k) g[k] = k < a.K ? a.gate[k] : 0.f;
  const long nvec = a.n >> 2;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < a.K) {
        float4 o = *reinterpret_cast<const float4*>(a.o[k] + 4 * v);
        acc.x = fmaf(g[k], o.x, acc.x); acc.y = fmaf(g[k], o.y, acc.y);
        acc.z = fmaf(g[k], o.z, acc.z); acc.w = fmaf(g[k], o.w, acc.w);
      }
    *reinterpret_cast<float4*>(a.out + 4 * v) = acc;
  }
}

// gate_grad[k] = <o_k, dout>;  d_o[k] = gate[k] * dout for the candidates that take gradient.
__global__ void __launch_bounds__(EW_THREADS) mixed_alpha_dot_kernel(MixedArgs a) {
  __shared__ float red[MAXK][EW_THREADS / 32];
  pdl_wait(); pdl_launch();
  float g[MAXK], dot[MAXK];
#pragma unroll
  for (int k = 0; k < MAXK; ++k) { g[k] = k < a.K ? a.gate[k] : 0.f; dot[k] = 0.f; }
  const long nvec = a.n >> 2;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
    const float4 d = *reinterpret_cast<const float4*>(a.dout + 4 * v);
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < a.K) {
        const float4 o = *reinterpret_cast<const float4*>(a.o[k] + 4 * v);
        dot[k] += (o.x * d.x + o.y * d.y) + (o.z * d.z + o.w * d.w);
        if (a.d_o[k]) {
          float4 r = make_float4(g[k] * d.x, g[k] * d.y, g[k] * d.z, g[k] * d.w);
          *reinterpret_cast<float4*>(a.d_o[k] + 4 * v) = r;
        }
      }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < MAXK; ++k) {
    float s = warp_sum(dot[k]);
    if (lane == 0) red[k][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < a.K) {
    float tot = 0.f;
    for (int w = 0; w < EW_THREADS / 32; ++w) tot += red[threadIdx.x][w];
    atomicAdd(&a.gate_grad[threadIdx.x], tot);
  }
}

// One launch casts every weight of the model: table[c] = {src pointer, dst pointer, element count (<= 4096, % 4 == 0)}
__global__ void __launch_bounds__(EW_THREADS) cast_multi_kernel(const long long* __restrict__ table) {
  pdl_wait(); pdl_launch();
  const long long* e = table + 3 * (long long)blockIdx.x;
  const float* src = reinterpret_cast<const float*>(e[0]);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e[1]);
  const int nvec = (int)(e[2] >> 2);
  for (int v = threadIdx.x; v < nvec; v += EW_THREADS) {
    const float4 f = *reinterpret_cast<const float4*>(src + 4 * v);
    __nv_bfloat162 lo = __floats2bfloat162_rn(f.x, f.y), hi = __floats2bfloat162_rn(f.z, f.w);
    uint2 pk;
    pk.x = *reinterpret_cast<unsigned*>(&lo);
    pk.y = *reinterpret_cast<unsigned*>(&hi);
    *reinterpret_cast<uint2*>(dst + 4 * v) = pk;
  }
}

// ---- stem: one pass over the region features gives their bf16 copy (A operand of the image
// projection GEMM) and the padding mask make_mask() derives from them (full_vqa.py:113-114: sum|x| == 0) -----------
__global__ void __launch_bounds__(EW_THREADS) cast_rowmask_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ x16,
                                                                  unsigned char* __restrict__ mask, int cols) {
  __shared__ int any_nz;
  pdl_wait(); pdl_launch();
  if (threadIdx.x == 0) any_nz = 0;
  __syncthreads();
  const long base = (long)blockIdx.x * cols;
  const int nvec = cols >> 2;
  bool nz = false;
  for (int v = threadIdx.x; v < nvec; v += EW_THREADS) {
    const float4 f = *reinterpret_cast<const float4*>(x + base + 4 * v);
    nz |= (f.x != 0.f) | (f.y != 0.f) | (f.z != 0.f) | (f.w != 0.f);
    if (x16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(f.x, f.y), hi = __floats2bfloat162_rn(f.z, f.w);
      uint2 pk;
      pk.x = *reinterpret_cast<unsigned*>(&lo);
      pk.y = *reinterpret_cast<unsigned*>(&hi);
      *reinterpret_cast<uint2*>(x16 + base + 4 * v) = pk;
    }
  }
  if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) atomicOr(&any_nz, 1);
  __syncthreads();
  if (threadIdx.x == 0) mask[blockIdx.x] = any_nz ? 0 : 1;
}

// make_mask() over a tensor that is already bf16 (a loader that ships bf16 region features): mask[r] = 1 iff row r is all zero
__global__ void __launch_bounds__(EW_THREADS) rowmask_bf16_kernel(const uint2* __restrict__ x, unsigned char* __restrict__ mask, int cols) {
  __shared__ int any_nz;
  pdl_wait(); pdl_launch();
  if (threadIdx.x == 0) any_nz = 0;
  __syncthreads();
  const uint2* row = x + (size_t)blockIdx.x * (cols >> 2);
  bool nz = false;
  for (int v = threadIdx.x; v < (cols >> 2); v += EW_THREADS) {
    const uint2 w = row[v];
    nz |= ((w.x | w.y) & 0x7FFF7FFFu) != 0u;          // +0 / -0 both count as zero, as in sum(|x|) == 0
  }
  if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) atomicOr(&any_nz, 1);
  __syncthreads();
  if (threadIdx.x == 0) mask[blockIdx.x] = any_nz ? 0 : 1;
}

// ---- geometry producer: the per-sample CPU step of the reference loader on the device ------------
// relation_embedding (load_data_vqa.py:7-33) + the zero padding of :236-239: for valid boxes i, j
//   g = (log max(|cx_i - cx_j| / w_i, 1e-3), log max(|cy_i - cy_j| / h_i, 1e-3), log(w_i / w_j), log(h_i / h_j)),
// w = x2 - x1 + 1, h = y2 - y1 + 1, c = (min + max) / 2; pairs that involve a padded region are all-zero.  float32
// arithmetic in the reference's operation order (IEEE divide, logf <= 1 ulp): the batch ships [B,N,4] boxes instead of
// the [B,N,N,4] tensor (10 MB per 64 samples less host->device traffic).
__global__ void __launch_bounds__(EW_THREADS) box_geometry_kernel(const float4* __restrict__ boxes, const unsigned char* __restrict__ pad,
                                                                  float4* __restrict__ g4, int N) {
  pdl_wait(); pdl_launch();
  const int b = blockIdx.y;
  const int nn = N * N;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += gridDim.x * blockDim.x) {
    const int i = e / N, j = e - i * N;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool valid = !pad || (pad[(size_t)b * N + i] == 0 && pad[(size_t)b * N + j] == 0);
    if (valid) {
      const float4 bi = __ldg(boxes + (size_t)b * N + i), bj = __ldg(boxes + (size_t)b * N + j);
      const float cxi = (bi.x + bi.z) * 0.5f, cyi = (bi.y + bi.w) * 0.5f, wi = (bi.z - bi.x) + 1.f, hi = (bi.w - bi.y) + 1.f;
      const float cxj = (bj.x + bj.z) * 0.5f, cyj = (bj.y + bj.w) * 0.5f, wj = (bj.z - bj.x) + 1.f, hj = (bj.w - bj.y) + 1.f;
      out.x = logf(fmaxf(fabsf(__fdiv_rn(cxi - cxj, wi)), 1e-3f));
      out.y = logf(fmaxf(fabsf(__fdiv_rn(cyi - cyj, hi)), 1e-3f));
      out.z = logf(__fdiv_rn(wi, wj));
      out.w = logf(__fdiv_rn(hi, hj));
    }
    g4[(size_t)b * nn + e] = out;
  }
}

// ---- optimizer tail: clip_grad_norm_ + Adam in two passes over flat / tabled buffers ----------
// Deterministic: every block stores its partial, the last block to finish adds them in block order.  (An atomicAdd
// into one float would make the clip coefficient depend on block scheduling, and data-parallel replicas that apply
// the same averaged gradient would drift apart bit by bit.)  scratch[0] = arrival counter, scratch[1..] = partials.
__global__ void __launch_bounds__(EW_THREADS) sumsq_kernel(const float* __restrict__ x, long n, float* __restrict__ out,
                                                           float* __restrict__ scratch) {
  __shared__ float red[EW_THREADS / 32];
  __shared__ bool last;
  pdl_wait(); pdl_launch();
  float s = 0.f;
  const long nvec = n >> 2;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long)gridDim.x * blockDim.x) {
    const float4 f = *reinterpret_cast<const float4*>(x + 4 * v);
    s += (f.x * f.x + f.y * f.y) + (f.z * f.z + f.w * f.w);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  unsigned* counter = reinterpret_cast<unsigned*>(scratch);
  float* partial = scratch + 1;
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < EW_THREADS / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
    __threadfence();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += EW_THREADS) t += __ldcg(partial + b);   // fixed assignment
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < EW_THREADS / 32; ++w) tot += red[w];
    *out = tot;
    *counter = 0u;
  }
}

// table[c] = {param*, grad*, exp_avg*, exp_avg_sq*, count (<= 4096)}.  torch.optim.Adam arithmetic
// (no amsgrad, no weight decay): p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps), with the gradient first
// scaled by clip_grad_norm_'s coefficient min(1, max_norm / (||g|| + 1e-6)).
__global__ void __launch_bounds__(EW_THREADS) clip_adam_kernel(const long long* __restrict__ table, const float* __restrict__ sumsq,
                                                               const float* __restrict__ lr_ptr, const unsigned long long* __restrict__ step_state,
                                                               float b1, float b2, float eps, float max_norm) {
  pdl_wait(); pdl_launch();
  const long long* e = table + 5 * (long long)blockIdx.x;
  float* p = reinterpret_cast<float*>(e[0]);
  const float* g = reinterpret_cast<const float*>(e[1]);
  float* m = reinterpret_cast<float*>(e[2]);
  float* v = reinterpret_cast<float*>(e[3]);
  const int count = (int)e[4];
  const bool aligned = (((e[0] | e[1] | e[2] | e[3]) & 15) == 0);
  const int nvec = aligned ? (count >> 2) : 0;
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(*sumsq) + 1e-6f));
  const float t = (float)step_state[1];
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = *lr_ptr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (int i = threadIdx.x; i < nvec; i += EW_THREADS) {
    float4 gv = *reinterpret_cast<const float4*>(g + 4 * i);
    float4 mv = *reinterpret_cast<const float4*>(m + 4 * i);
    float4 vv = *reinterpret_cast<const float4*>(v + 4 * i);
    float4 pv = *reinterpret_cast<const float4*>(p + 4 * i);
    float gg[4] = {gv.x * coef, gv.y * coef, gv.z * coef, gv.w * coef};
    float mm[4] = {mv.x, mv.y, mv.z, mv.w}, vq[4] = {vv.x, vv.y, vv.z, vv.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      mm[u] = b1 * mm[u] + (1.f - b1) * gg[u];
      vq[u] = b2 * vq[u] + (1.f - b2) * gg[u] * gg[u];
      pp[u] -= step_size * mm[u] / (sqrtf(vq[u]) * inv_sqrt_bc2 + eps);
    }
    *reinterpret_cast<float4*>(m + 4 * i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + 4 * i) = make_float4(vq[0], vq[1], vq[2], vq[3]);
    *reinterpret_cast<float4*>(p + 4 * i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
  }
  for (int i = 4 * nvec + threadIdx.x; i < count; i += EW_THREADS) {     // unaligned tensors and tails
    const float gg = g[i] * coef;
    const float mm = b1 * m[i] + (1.f - b1) * gg;
    const float vq = b2 * v[i] + (1.f - b2) * gg * gg;
    m[i] = mm; v[i] = vq;
    p[i] -= step_size * mm / (sqrtf(vq) * inv_sqrt_bc2 + eps);
  }
}

__global__ void rng_advance_kernel(unsigned long long* state) { state[1] += 1ull; }

}  // namespace

extern "C" int mmnas_cast_f32_to_bf16(const float* src, void* dst, long n, mmnas_stream stream) {
  MMNAS_CHECK_ARG(n >= 0, "cast: negative length");
  if (n == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(src && dst, "cast: null buffer");
  MMNAS_CHECK_ARG(((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 8) == 0, "cast: buffers must be 16-byte aligned");
  MMNAS_CUDA(mmnas_launch(cast_kernel, dim3(ew_grid(n >> 2)), dim3(EW_THREADS), 0, (cudaStream_t)stream, src, (__nv_bfloat16*)dst, n));
  return MMNAS_OK;
}

extern "C" int mmnas_cast_multi(const void* table, int n_chunks, mmnas_stream stream) {
  MMNAS_CHECK_ARG(n_chunks >= 0, "cast_multi: negative chunk count");
  if (n_chunks == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(table, "cast_multi: null table");
  MMNAS_CUDA(mmnas_launch(cast_multi_kernel, dim3(n_chunks), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const long long*)table));
  return MMNAS_OK;
}

extern "C" int mmnas_colsum(int dtype, const void* x, int rows, int cols, long ld, float* out, int accumulate,
                            mmnas_stream stream) {
  MMNAS_CHECK_ARG(dtype == 0 || dtype == 1, "colsum: dtype");
  MMNAS_CHECK_ARG(rows >= 0 && cols > 0 && x && out, "colsum: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (!accumulate) MMNAS_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, s));
  if (rows == 0) return MMNAS_OK;
  const int gx = ceil_div(cols, 256);
  int gy = ceil_div(rows, 8 * 4);
  const int cap = gx >= 8 ? 40 : (gx >= 2 ? 148 : 296);     // ~2 CTAs per SM in total
  if (gy > cap) gy = cap;
  MMNAS_CHECK_ARG(((uintptr_t)x % 16) == 0, "colsum: x must be 16-byte aligned");
  dim3 grid(gx, gy);
  if (dtype == 0) MMNAS_CUDA(mmnas_launch(colsum_kernel<float>, grid, dim3(256), 0, s, (const float*)x, rows, cols, ld, out));
  else MMNAS_CUDA(mmnas_launch(colsum_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, (const __nv_bfloat16*)x, rows, cols, ld, out));
  return MMNAS_OK;
}

extern "C" int mmnas_mixed_accum(int K, const float* const* outs, const float* gate, float* out, long n,
                                 mmnas_stream stream) {
  MMNAS_CHECK_ARG(K >= 1 && K <= MAXK, "mixed_accum: 1..8 candidates");
  MMNAS_CHECK_ARG(outs && gate && out && n >= 0 && (n % 4) == 0, "mixed_accum: bad argument (n must be a multiple of 4)");
  if (n == 0) return MMNAS_OK;
  MixedArgs a = {};
  a.K = K; a.gate = gate; a.out = out; a.n = n;
  for (int k = 0; k < K; ++k) { MMNAS_CHECK_ARG(outs[k], "mixed_accum: null candidate output"); a.o[k] = outs[k]; }
  MMNAS_CUDA(mmnas_launch(mixed_accum_kernel, dim3(ew_grid(n >> 2)), dim3(EW_THREADS), 0, (cudaStream_t)stream, a));
  return MMNAS_OK;
}

extern "C" int mmnas_mixed_alpha_dot(int K, const float* const* outs, const float* gate, const float* dout,
                                     float* gate_grad, float* const* d_outs, long n, mmnas_stream stream) {
  MMNAS_CHECK_ARG(K >= 1 && K <= MAXK, "mixed_alpha_dot: 1..8 candidates");
  MMNAS_CHECK_ARG(outs && gate && dout && gate_grad && n >= 0 && (n % 4) == 0, "mixed_alpha_dot: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  MMNAS_CUDA(cudaMemsetAsync(gate_grad, 0, sizeof(float) * K, s));
  if (n == 0) return MMNAS_OK;
  MixedArgs a = {};
  a.K = K; a.gate = gate; a.dout = dout; a.gate_grad = gate_grad; a.n = n;
  for (int k = 0; k < K; ++k) {
    MMNAS_CHECK_ARG(outs[k], "mixed_alpha_dot: null candidate output");
    a.o[k] = outs[k];
    a.d_o[k] = d_outs ? d_outs[k] : nullptr;
  }
  MMNAS_CUDA(mmnas_launch(mixed_alpha_dot_kernel, dim3(ew_grid(n >> 2)), dim3(EW_THREADS), 0, s, a));
  return MMNAS_OK;
}

extern "C" int mmnas_cast_rowmask(const float* x, void* x_bf16, unsigned char* mask, int rows, int cols, mmnas_stream stream) {
  MMNAS_CHECK_ARG(rows >= 0 && cols > 0 && (cols % 4) == 0, "cast_rowmask: cols must be a multiple of 4");
  if (rows == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(x && mask && ((uintptr_t)x % 16) == 0 && ((uintptr_t)x_bf16 % 8) == 0, "cast_rowmask: null / misaligned buffer");
  MMNAS_CUDA(mmnas_launch(cast_rowmask_kernel, dim3(rows), dim3(EW_THREADS), 0, (cudaStream_t)stream, x, (__nv_bfloat16*)x_bf16, mask, cols));
  return MMNAS_OK;
}

extern "C" int mmnas_rowmask_bf16(const void* x_bf16, unsigned char* mask, int rows, int cols, mmnas_stream stream) {
  MMNAS_CHECK_ARG(rows >= 0 && cols > 0 && (cols % 4) == 0, "rowmask_bf16: cols must be a multiple of 4");
  if (rows == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(x_bf16 && mask && ((uintptr_t)x_bf16 % 8) == 0, "rowmask_bf16: null / misaligned buffer");
  MMNAS_CUDA(mmnas_launch(rowmask_bf16_kernel, dim3(rows), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const uint2*)x_bf16, mask, cols));
  return MMNAS_OK;
}

extern "C" int mmnas_box_geometry(const float* boxes, const unsigned char* pad_mask, float* g4, int B, int N, mmnas_stream stream) {
  MMNAS_CHECK_ARG(B >= 0 && N >= 1, "box_geometry: bad sizes");
  if (B == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(boxes && g4 && ((uintptr_t)boxes % 16) == 0 && ((uintptr_t)g4 % 16) == 0, "box_geometry: null / misaligned buffer");
  const int gx = ceil_div(N * N, EW_THREADS);
  MMNAS_CUDA(mmnas_launch(box_geometry_kernel, dim3(gx, B), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const float4*)boxes, pad_mask,
                          (float4*)g4, N));
  return MMNAS_OK;
}

extern "C" int mmnas_sumsq_f32(const float* x, long n, float* out, float* scratch, mmnas_stream stream) {
  MMNAS_CHECK_ARG(out && scratch && n >= 0 && (n % 4) == 0 && (n == 0 || (x && ((uintptr_t)x % 16) == 0)),
                  "sumsq: bad argument (null buffer, n % 4, 16-byte alignment)");
  static_assert(148 * 8 + 1 <= MMNAS_SUMSQ_SCRATCH, "scratch too small for the grid cap");
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) { MMNAS_CUDA(cudaMemsetAsync(out, 0, sizeof(float), s)); return MMNAS_OK; }
  MMNAS_CUDA(cudaMemsetAsync(scratch, 0, sizeof(unsigned), s));
  MMNAS_CUDA(mmnas_launch(sumsq_kernel, dim3(ew_grid(n >> 2)), dim3(EW_THREADS), 0, s, x, n, out, scratch));
  return MMNAS_OK;
}

extern "C" int mmnas_clip_adam(const void* table, int n_chunks, const float* sumsq, const float* lr,
                               const unsigned long long* step_state, float beta1, float beta2, float eps, float max_norm,
                               mmnas_stream stream) {
  MMNAS_CHECK_ARG(n_chunks >= 0, "clip_adam: negative chunk count");
  if (n_chunks == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(table && lr && step_state && (max_norm <= 0.f || sumsq), "clip_adam: null argument");
  MMNAS_CUDA(mmnas_launch(clip_adam_kernel, dim3(n_chunks), dim3(EW_THREADS), 0, (cudaStream_t)stream,
                          (const long long*)table, sumsq, lr, step_state, beta1, beta2, eps, max_norm));
  return MMNAS_OK;
}

extern "C" int mmnas_rng_advance(unsigned long long* state, mmnas_stream stream) {
  MMNAS_CHECK_ARG(state, "rng_advance: null state");
  mmnas_count_launch();
  rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
  MMNAS_LAUNCH_CHECK();
  return MMNAS_OK;
}

static std::atomic<unsigned long long> g_launches{0};
void mmnas_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long mmnas_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

bool mmnas_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MMNAS_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// ---- error string / ABI version ------------------------------------------------------------
static thread_local char g_err[256] = "";
void mmnas_set_error(const char* msg) {
  int i = 0;
  for (; msg && msg[i] && i < 255; ++i) g_err[i] = msg[i];
  g_err[i] = 0;
}
extern "C" const char* mmnas_last_error(void) { return g_err; }
extern "C" int mmnas_abi_version(void) { return MMNAS_B200_ABI_VERSION; }
