// Multi-head attention core for the short padded MMnas sequences (Nq, Nk <= 128, head dim 64).
//   S = Q K^T * scale (+ bias);  S[mask] = -1e9;  P = softmax(S);  P = dropout(P);  O = P V
// Reference: MHAtt.att modules.py:190-199 and RelMHAtt.forward :232-240.  masked_fill REPLACES the
// logit after the bias add, so a fully padded key row yields a uniform distribution, not NaN.
// One CTA per (sample, head): Q, K, V live in shared memory, logits stay in registers (a warp owns
// 4 query rows, lane l owns keys l, l+32, l+64, l+96), nothing but O reaches HBM.  The backward
// recomputes P from Q, K (+bias, mask, dropout key) instead of reading a saved [B,h,Nq,Nk] map.
// fp32 math throughout; T (float / bf16) is only the storage type of q, k, v, o and their grads.
#include <cstdlib>
#include "common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int D = 64;        // head dim ('*_64' ops)
constexpr int NWARP = 8;
constexpr int RPW = 4;       // query rows per warp per pass
constexpr int KB = 4;        // key blocks of 32 per lane -> Nk <= 128

struct AttnArgs {
  int B, heads, Nq, Nk;
  const void *q, *k, *v; long ldq, ldk, ldv;
  const unsigned char* kmask;   // [B, Nk], 1 = padded key
  const float* bias;            // [B, heads, Nq, Nk] or null
  void* o; long ldo;
  float scale;
  DropCfg drop;
  // backward only
  const void* dout; long lddo;
  void *dq, *dk, *dv; long lddq, lddk, lddv;
  float* dbias;
};

template <typename T>
__device__ __forceinline__ void load_tile(float* dst, int dst_ld, const T* src, long ld, int rows, int tid, int nthr) {
  // rows x 64, coalesced along the head dim
  for (int e = tid; e < rows * D; e += nthr) {
    int r = e / D, c = e % D;
    dst[r * dst_ld + c] = to_f32<T>(src[(long)r * ld + c]);
  }
}

// logits + softmax for the 4 rows [i0, i0+4) of this warp.  On exit p[a][b] holds the softmax
// probabilities (before dropout) for key j = lane + 32 b; if WITH_DP, dp[a][b] = dO_i . V_j.
template <bool WITH_DP>
__device__ __forceinline__ void rows_softmax(const AttnArgs& a, int b, int h, int i0, int lane, const float* Qs,
                                             const float* Ks, const float* dOs, const float* Vs,
                                             float (&p)[RPW][KB], float (&dp)[RPW][KB]) {
  const int Nq = a.Nq, Nk = a.Nk;
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) { p[r][kb] = 0.f; dp[r][kb] = 0.f; }
#pragma unroll 4
  for (int c = 0; c < D; ++c) {
    float qv[RPW], dv[RPW], kv[KB], vv[KB];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      int i = min(i0 + r, Nq - 1);
      qv[r] = Qs[i * D + c];
      if (WITH_DP) dv[r] = dOs[i * D + c];
    }
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      int j = min(lane + 32 * kb, Nk - 1);
      kv[kb] = Ks[j * (D + 1) + c];
      if (WITH_DP) vv[kb] = Vs[j * (D + 1) + c];
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        p[r][kb] = fmaf(qv[r], kv[kb], p[r][kb]);
        if (WITH_DP) dp[r][kb] = fmaf(dv[r], vv[kb], dp[r][kb]);
      }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int i = i0 + r;
    float mx = -INFINITY;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      const int j = lane + 32 * kb;
      float s = -INFINITY;
      if (j < Nk && i < Nq) {
        s = p[r][kb] * a.scale;
        if (a.bias) s += a.bias[(((long)b * a.heads + h) * Nq + i) * Nk + j];
        if (a.kmask && a.kmask[(long)b * Nk + j]) s = -1e9f;
      }
      p[r][kb] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      float e = (p[r][kb] == -INFINITY) ? 0.f : expf(p[r][kb] - mx);
      p[r][kb] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) p[r][kb] *= inv;
  }
}

template <typename T>
__global__ void __launch_bounds__(NWARP * 32) attn_fwd_kernel(AttnArgs a) {
  extern __shared__ float sm[];
  const int h = blockIdx.x, b = blockIdx.y;
  const int Nq = a.Nq, Nk = a.Nk;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NkP = (Nk + 3) & ~3;
  float* Qs = sm;                        // [Nq][64]
  float* Ks = Qs + Nq * D;               // [Nk][65]
  float* Vs = Ks + Nk * (D + 1);         // [Nk][64]
  float* Ps = Vs + Nk * D;               // [NWARP][RPW][NkP]
  const T* q = (const T*)a.q + (long)b * Nq * a.ldq + h * D;
  const T* k = (const T*)a.k + (long)b * Nk * a.ldk + h * D;
  const T* v = (const T*)a.v + (long)b * Nk * a.ldv + h * D;
  load_tile<T>(Qs, D, q, a.ldq, Nq, tid, blockDim.x);
  load_tile<T>(Ks, D + 1, k, a.ldk, Nk, tid, blockDim.x);
  load_tile<T>(Vs, D, v, a.ldv, Nk, tid, blockDim.x);
  __syncthreads();
  const bool use_drop = a.drop.state != nullptr && a.drop.thresh > 0;
  const uint64_t key = use_drop ? drop_key(a.drop) : 0;
  float* Pw = Ps + warp * RPW * NkP;
  T* o = (T*)a.o + (long)b * Nq * a.ldo + h * D;
  for (int i0 = warp * RPW; i0 < Nq; i0 += NWARP * RPW) {
    float p[RPW][KB], dp[RPW][KB];
    rows_softmax<false>(a, b, h, i0, lane, Qs, Ks, nullptr, nullptr, p, dp);
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const int j = lane + 32 * kb, i = i0 + r;
        if (j < Nk) {
          float pv = p[r][kb];
          if (use_drop && i < Nq)
            pv *= drop_mult(key, (((uint64_t)b * a.heads + h) * Nq + i) * Nk + j, a.drop.thresh, a.drop.scale);
          Pw[r * NkP + j] = pv;
        }
      }
    __syncwarp();
    float acc[RPW][2];
#pragma unroll
    for (int r = 0; r < RPW; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
    for (int j = 0; j < Nk; ++j) {
      const float v0 = Vs[j * D + lane], v1 = Vs[j * D + lane + 32];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const float pv = Pw[r * NkP + j];
        acc[r][0] = fmaf(pv, v0, acc[r][0]);
        acc[r][1] = fmaf(pv, v1, acc[r][1]);
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int i = i0 + r;
      if (i < Nq) {
        o[(long)i * a.ldo + lane] = from_f32<T>(acc[r][0]);
        o[(long)i * a.ldo + lane + 32] = from_f32<T>(acc[r][1]);
      }
    }
    __syncwarp();
  }
}

// Backward: dV = Pd^T dO;  dPd = dO V^T;  dS = P * (m*dPd - delta),  delta_i = dO_i . O_i;
//           dS[mask] = 0 (masked_fill cuts the graph);  dbias = dS;  dQ = scale dS K;  dK = scale dS^T Q.
template <typename T>
__global__ void __launch_bounds__(NWARP * 32) attn_bwd_kernel(AttnArgs a) {
  extern __shared__ float sm[];
  const int h = blockIdx.x, b = blockIdx.y;
  const int Nq = a.Nq, Nk = a.Nk;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NkP = Nk + 1;
  float* Qs = sm;                          // [Nq][64]
  float* dOs = Qs + Nq * D;                // [Nq][64]
  float* Ks = dOs + Nq * D;                // [Nk][65]
  float* Vs = Ks + Nk * (D + 1);           // [Nk][65]
  float* Ps = Vs + Nk * (D + 1);           // [Nq][NkP]  dropped-out probabilities
  float* dSs = Ps + Nq * NkP;              // [Nq][NkP]
  const T* q = (const T*)a.q + (long)b * Nq * a.ldq + h * D;
  const T* k = (const T*)a.k + (long)b * Nk * a.ldk + h * D;
  const T* v = (const T*)a.v + (long)b * Nk * a.ldv + h * D;
  const T* o = (const T*)a.o + (long)b * Nq * a.ldo + h * D;
  const T* dO = (const T*)a.dout + (long)b * Nq * a.lddo + h * D;
  load_tile<T>(Qs, D, q, a.ldq, Nq, tid, blockDim.x);
  load_tile<T>(dOs, D, dO, a.lddo, Nq, tid, blockDim.x);
  load_tile<T>(Ks, D + 1, k, a.ldk, Nk, tid, blockDim.x);
  load_tile<T>(Vs, D + 1, v, a.ldv, Nk, tid, blockDim.x);
  __syncthreads();
  const bool use_drop = a.drop.state != nullptr && a.drop.thresh > 0;
  const uint64_t key = use_drop ? drop_key(a.drop) : 0;
  T* dq = (T*)a.dq + (long)b * Nq * a.lddq + h * D;
  for (int i0 = warp * RPW; i0 < Nq; i0 += NWARP * RPW) {
    float p[RPW][KB], dp[RPW][KB];
    rows_softmax<true>(a, b, h, i0, lane, Qs, Ks, dOs, Vs, p, dp);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int i = i0 + r;
      if (i >= Nq) continue;   // warp-uniform
      // delta_i = dO_i . O_i   (O read from global: it is the saved forward output)
      float dl = dOs[i * D + lane] * to_f32<T>(o[(long)i * a.ldo + lane]) +
                 dOs[i * D + lane + 32] * to_f32<T>(o[(long)i * a.ldo + lane + 32]);
      dl = warp_sum(dl);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const int j = lane + 32 * kb;
        if (j < Nk) {
          float m = 1.f;
          if (use_drop) m = drop_mult(key, (((uint64_t)b * a.heads + h) * Nq + i) * Nk + j, a.drop.thresh, a.drop.scale);
          float ds = p[r][kb] * (m * dp[r][kb] - dl);
          if (a.kmask && a.kmask[(long)b * Nk + j]) ds = 0.f;
          Ps[i * NkP + j] = p[r][kb] * m;
          dSs[i * NkP + j] = ds;
          if (a.dbias) a.dbias[(((long)b * a.heads + h) * Nq + i) * Nk + j] = ds;
        }
      }
    }
    __syncwarp();
    // dQ rows of this warp
    float acc[RPW][2];
#pragma unroll
    for (int r = 0; r < RPW; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
    for (int j = 0; j < Nk; ++j) {
      const float k0 = Ks[j * (D + 1) + lane], k1 = Ks[j * (D + 1) + lane + 32];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const float ds = (i0 + r < Nq) ? dSs[(i0 + r) * NkP + j] : 0.f;
        acc[r][0] = fmaf(ds, k0, acc[r][0]);
        acc[r][1] = fmaf(ds, k1, acc[r][1]);
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int i = i0 + r;
      if (i < Nq) {
        dq[(long)i * a.lddq + lane] = from_f32<T>(acc[r][0] * a.scale);
        dq[(long)i * a.lddq + lane + 32] = from_f32<T>(acc[r][1] * a.scale);
      }
    }
  }
  __syncthreads();
  // dK, dV: a warp owns 4 key rows, lanes own head-dim columns lane, lane+32
  T* dk = (T*)a.dk + (long)b * Nk * a.lddk + h * D;
  T* dv = (T*)a.dv + (long)b * Nk * a.lddv + h * D;
  for (int j0 = warp * RPW; j0 < Nk; j0 += NWARP * RPW) {
    float ak[RPW][2], av[RPW][2];
#pragma unroll
    for (int r = 0; r < RPW; ++r) { ak[r][0] = ak[r][1] = av[r][0] = av[r][1] = 0.f; }
    for (int i = 0; i < Nq; ++i) {
      const float q0 = Qs[i * D + lane], q1 = Qs[i * D + lane + 32];
      const float d0 = dOs[i * D + lane], d1 = dOs[i * D + lane + 32];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const int j = min(j0 + r, Nk - 1);
        const float ds = dSs[i * NkP + j], pd = Ps[i * NkP + j];
        ak[r][0] = fmaf(ds, q0, ak[r][0]); ak[r][1] = fmaf(ds, q1, ak[r][1]);
        av[r][0] = fmaf(pd, d0, av[r][0]); av[r][1] = fmaf(pd, d1, av[r][1]);
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int j = j0 + r;
      if (j < Nk) {
        dk[(long)j * a.lddk + lane] = from_f32<T>(ak[r][0] * a.scale);
        dk[(long)j * a.lddk + lane + 32] = from_f32<T>(ak[r][1] * a.scale);
        dv[(long)j * a.lddv + lane] = from_f32<T>(av[r][0]);
        dv[(long)j * a.lddv + lane + 32] = from_f32<T>(av[r][1]);
      }
    }
  }
}

size_t fwd_smem(int Nq, int Nk) {
  int NkP = (Nk + 3) & ~3;
  return sizeof(float) * ((size_t)Nq * D + (size_t)Nk * (D + 1) + (size_t)Nk * D + (size_t)NWARP * RPW * NkP);
}
size_t bwd_smem(int Nq, int Nk) {
  return sizeof(float) * (2 * (size_t)Nq * D + 2 * (size_t)Nk * (D + 1) + 2 * (size_t)Nq * (Nk + 1));
}

int check_common(int dtype, int B, int heads, int Nq, int Nk, int head_dim) {
  MMNAS_CHECK_ARG(dtype == 0 || dtype == 1, "attn: dtype must be 0 (f32) or 1 (bf16)");
  MMNAS_CHECK_ARG(head_dim == D, "attn: only head dim 64 (the '*_64' operators) is implemented");
  MMNAS_CHECK_ARG(B >= 0 && heads > 0 && Nq > 0 && Nk > 0, "attn: bad sizes");
  MMNAS_CHECK_ARG(Nk <= 32 * KB, "attn: Nk > 128 not supported");
  return MMNAS_OK;
}

}  // namespace

// tensor-core implementations (attention_tc.cu); MMNAS_ERR_UNSUPPORTED = operands miss the TMA constraints
int mmnas_attn_fwd_tc(int B, int heads, int Nq, int Nk, const void* q, long ldq, const void* k, long ldk, const void* v,
                      long ldv, const unsigned char* kmask, const float* bias, void* o, long ldo, float scale,
                      const unsigned long long* rng_state, unsigned long long salt, float p, cudaStream_t s);
int mmnas_attn_bwd_tc(int B, int heads, int Nq, int Nk, const void* q, long ldq, const void* k, long ldk, const void* v,
                      long ldv, const unsigned char* kmask, const float* bias, const void* o, long ldo, const void* dout,
                      long lddo, void* dq, long lddq, void* dk, long lddk, void* dv, long lddv, float* dbias, float scale,
                      const unsigned long long* rng_state, unsigned long long salt, float p, cudaStream_t s);

// tuning only (MMNAS_ATTN_TC_MIN_NK): key counts below this take the FFMA kernel in the bf16 arm too
static int tc_min_nk() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MMNAS_ATTN_TC_MIN_NK"); v = e ? atoi(e) : 0; }
  return v;
}

extern "C" int mmnas_attn_fwd(int dtype, int B, int heads, int Nq, int Nk, int head_dim, const void* q, long ldq,
                              const void* k, long ldk, const void* v, long ldv, const unsigned char* kmask,
                              const float* bias, void* o, long ldo, float scale, const unsigned long long* rng_state,
                              unsigned long long salt, float p, mmnas_stream stream) {
  int rc = check_common(dtype, B, heads, Nq, Nk, head_dim);
  if (rc) return rc;
  if (B == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(q && k && v && o, "attn_fwd: null operand");
  if (dtype == 1 && Nk >= tc_min_nk()) {   // bf16 arm: tcgen05 kernel; the FFMA kernel below only if TMA cannot address the operands
    rc = mmnas_attn_fwd_tc(B, heads, Nq, Nk, q, ldq, k, ldk, v, ldv, kmask, bias, o, ldo, scale, rng_state, salt, p,
                           (cudaStream_t)stream);
    if (rc != MMNAS_ERR_UNSUPPORTED) return rc;
  }
  size_t smem = fwd_smem(Nq, Nk);
  MMNAS_CHECK_ARG(smem <= 227 * 1024, "attn_fwd: sequence too long for the shared-memory tile");
  AttnArgs a = {};
  a.B = B; a.heads = heads; a.Nq = Nq; a.Nk = Nk;
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.kmask = kmask; a.bias = bias; a.o = o; a.ldo = ldo; a.scale = scale;
  a.drop.state = p > 0.f ? rng_state : nullptr; a.drop.salt = salt;
  a.drop.thresh = (unsigned)(p * 65536.f + 0.5f); a.drop.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  dim3 grid(heads, B);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == 0) {
    MMNAS_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_kernel<float><<<grid, NWARP * 32, smem, s>>>(a);
  } else {
    MMNAS_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_kernel<__nv_bfloat16><<<grid, NWARP * 32, smem, s>>>(a);
  }
  MMNAS_LAUNCH_CHECK();
  return MMNAS_OK;
}

extern "C" int mmnas_attn_bwd(int dtype, int B, int heads, int Nq, int Nk, int head_dim, const void* q, long ldq,
                              const void* k, long ldk, const void* v, long ldv, const unsigned char* kmask,
                              const float* bias, const void* o, long ldo, const void* dout, long lddo, void* dq,
                              long lddq, void* dk, long lddk, void* dv, long lddv, float* dbias, float scale,
                              const unsigned long long* rng_state, unsigned long long salt, float p,
                              mmnas_stream stream) {
  int rc = check_common(dtype, B, heads, Nq, Nk, head_dim);
  if (rc) return rc;
  if (B == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(q && k && v && o && dout && dq && dk && dv, "attn_bwd: null operand");
  if (dtype == 1 && Nk >= tc_min_nk()) {
    rc = mmnas_attn_bwd_tc(B, heads, Nq, Nk, q, ldq, k, ldk, v, ldv, kmask, bias, o, ldo, dout, lddo, dq, lddq, dk, lddk,
                           dv, lddv, dbias, scale, rng_state, salt, p, (cudaStream_t)stream);
    if (rc != MMNAS_ERR_UNSUPPORTED) return rc;
  }
  size_t smem = bwd_smem(Nq, Nk);
  MMNAS_CHECK_ARG(smem <= 227 * 1024, "attn_bwd: sequence too long for the shared-memory tile");
  AttnArgs a = {};
  a.B = B; a.heads = heads; a.Nq = Nq; a.Nk = Nk;
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.kmask = kmask; a.bias = bias; a.o = (void*)o; a.ldo = ldo; a.scale = scale;
  a.dout = dout; a.lddo = lddo; a.dq = dq; a.dk = dk; a.dv = dv; a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
  a.dbias = dbias;
  a.drop.state = p > 0.f ? rng_state : nullptr; a.drop.salt = salt;
  a.drop.thresh = (unsigned)(p * 65536.f + 0.5f); a.drop.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  dim3 grid(heads, B);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == 0) {
    MMNAS_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_kernel<float><<<grid, NWARP * 32, smem, s>>>(a);
  } else {
    MMNAS_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_kernel<__nv_bfloat16><<<grid, NWARP * 32, smem, s>>>(a);
  }
  MMNAS_LAUNCH_CHECK();
  return MMNAS_OK;
}
