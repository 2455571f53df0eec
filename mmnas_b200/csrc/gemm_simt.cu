// fp32 SIMT GEMM: the full-precision arm of the hot path (parity gate: 1e-5 normwise against the
// reference's fp32 nn.Linear / torch.matmul, modules.py:18,38,172-175,193,199).  FFMA only, fp32
// accumulate in K order.  Generic strides cover forward (X W^T), dgrad (dY W) and wgrad (dY^T X)
// without transposed copies.  The bf16 tensor-core arm lives in gemm_tc.cu.
#include "common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

struct GemmArgs {
  int M, N, K;
  const float* A; long a_rs, a_cs;
  const float* B; long b_rs, b_cs;
  float* C; long ldc;
  const float* bias;
  int epilogue, accumulate;
  const float* aux; long ld_aux; float aux_scale;
  DropCfg drop;
};

// A_KCONTIG: A's k index is the unit-stride one (row-major [M,K]); else m is.  Same for B with n.
template <bool A_KCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // two-level accumulation: `part` sums 128 consecutive k, then is flushed into `acc`, so the round-off of a
  // long reduction (wgrad: K = tokens = 6400) stays at blocked-GEMM level instead of a serial 6400-term sum
  float acc[TM][TN], part[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) { acc[i][j] = 0.f; part[i][j] = 0.f; }

  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int r = 0; r < (BM * BK) / 256; ++r) {
      int idx = tid + r * 256;
      int mm = A_KCONTIG ? idx / BK : idx % BM;
      int kk = A_KCONTIG ? idx % BK : idx / BM;
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < g.M && gk < g.K) ? g.A[(long)gm * g.a_rs + (long)gk * g.a_cs] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < (BN * BK) / 256; ++r) {
      int idx = tid + r * 256;
      int nn = B_NCONTIG ? idx % BN : idx / BK;
      int kk = B_NCONTIG ? idx / BN : idx % BK;
      int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < g.N && gk < g.K) ? g.B[(long)gk * g.b_rs + (long)gn * g.b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
      b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
    if (((k0 / BK) & 7) == 7 || k0 + BK >= g.K) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) { acc[i][j] += part[i][j]; part[i][j] = 0.f; }
    }
    __syncthreads();
  }

  uint64_t key = 0;
  if (g.epilogue == 2) key = drop_key(g.drop);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int gm = m0 + ty * TM + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int gn = n0 + tx * TN + j;
      if (gn >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[gn];
      if (g.epilogue == 1 || g.epilogue == 2) v = fmaxf(v, 0.f);
      if (g.epilogue == 2) v *= drop_mult(key, (uint64_t)gm * g.N + gn, g.drop.thresh, g.drop.scale);
      if (g.epilogue == 3) v = g.aux[(long)gm * g.ld_aux + gn] > 0.f ? v * g.aux_scale : 0.f;
      float* c = g.C + (long)gm * g.ldc + gn;
      *c = g.accumulate ? *c + v : v;
    }
  }
}

}  // namespace

extern "C" int mmnas_gemm_f32(int M, int N, int K, const float* A, long a_rs, long a_cs, const float* B, long b_rs,
                              long b_cs, float* C, long ldc, const float* bias, int epilogue, int accumulate,
                              const float* aux, long ld_aux, float aux_scale, const unsigned long long* rng_state,
                              unsigned long long salt, float p, mmnas_stream stream) {
  MMNAS_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "gemm_f32: negative size");
  if (M == 0 || N == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(A && B && C, "gemm_f32: null operand");
  MMNAS_CHECK_ARG(epilogue >= 0 && epilogue <= 3, "gemm_f32: unknown epilogue");
  MMNAS_CHECK_ARG(epilogue != 3 || aux, "gemm_f32: epilogue 3 needs aux");
  if (epilogue == 2 && (p <= 0.f || !rng_state)) epilogue = 1;
  GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.a_rs = a_rs; g.a_cs = a_cs;
  g.B = B; g.b_rs = b_rs; g.b_cs = b_cs;
  g.C = C; g.ldc = ldc; g.bias = bias;
  g.epilogue = epilogue; g.accumulate = accumulate;
  g.aux = aux; g.ld_aux = ld_aux; g.aux_scale = aux_scale;
  g.drop.state = rng_state; g.drop.salt = salt;
  g.drop.thresh = (unsigned)(p * 65536.f + 0.5f);
  g.drop.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
  cudaStream_t s = (cudaStream_t)stream;
  const bool ak = (a_cs == 1), bn = (b_cs == 1);
  if (ak && bn) gemm_f32_kernel<true, true><<<grid, 256, 0, s>>>(g);
  else if (ak && !bn) gemm_f32_kernel<true, false><<<grid, 256, 0, s>>>(g);
  else if (!ak && bn) gemm_f32_kernel<false, true><<<grid, 256, 0, s>>>(g);
  else gemm_f32_kernel<false, false><<<grid, 256, 0, s>>>(g);
  MMNAS_LAUNCH_CHECK();
  return MMNAS_OK;
}
