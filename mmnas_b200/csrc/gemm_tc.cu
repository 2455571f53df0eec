// bf16 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory ->
// tcgen05.mma (cta_group::1, 128 x BN x 16, kind::f16, fp32 accumulators in TMEM) -> tcgen05.ld epilogue.
// Serves every dense contraction of the MMnas blocks in bf16 mode:
//   forward   Y = X W^T        A K-major  [M,K],  B K-major  [N,K]            (nn.Linear, modules.py:18,38,172-175)
//   dgrad     dX = dY W        A K-major  [M,N'], B MN-major (W itself, [N',K'])
//   wgrad     dW = dY^T X      A MN-major (dY, [M,N']), B MN-major (X, [M,K'])  + split-K over tokens
// so no transposed copy of an activation or a weight is ever materialised.
//
// v2: persistent, warp-specialised, software-pipelined over tiles.
//   grid = min(#work units, #SMs); a work unit = (split, m-tile, n-tile), n fastest so concurrently running CTAs
//   share the A tile in L2.  Warp 0 = TMA producer (STAGES-deep smem ring), warp 1 = MMA issuer (one elected lane)
//   + TMEM allocator, warps 2..9 = epilogue (two warps per TMEM lane quarter, alternating column chunks).  TMEM holds TWO accumulator buffers, so the epilogue of unit i
//   overlaps the main loop of unit i+1 (v1 paid prologue + fill + epilogue serially per tile and ran at 10 %).
//   Epilogue: tcgen05.ld (each warp owns the 32 TMEM lanes of its warp%4 quarter) -> padded smem staging ->
//   re-read row-contiguous, so every global access is a full 64/128-byte row segment; all fused epilogue math
//   (+bias, ReLU, dropout with one hash per 4 elements, ReLU-mask of a saved activation, fp32 accumulate,
//   bf16/fp32 store, split-K red.add.v4) happens in that coalesced layout.
//   BN = 256 (single 128x256x16 UMMA, half the operand traffic per FLOP) when it does not cost a wave.
#include <cstdlib>
#include "tc_common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 320;                      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int EPI_WARPS = 8;
constexpr int STG_PITCH = 36;                         // floats per staged row (32 + 4 pad: 16B aligned, conflict free)
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_PITCH * 4;   // one 32-row staging block per epilogue warp

template <int BN> struct Cfg {
  static constexpr int STAGES = BN == 128 ? 5 : 3;
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr uint32_t TMEM_COLS = 2 * BN;       // two accumulator buffers
};

struct TcEpilogue {
  int M, N, K;
  void* C; long ldc; int out_bf16;
  const float* bias;
  int relu;
  int accumulate;                    // fp32 out: C += result (plain RMW; split_k == 1)
  const __nv_bfloat16* aux; long ld_aux; float aux_scale;   // C = aux > 0 ? v * aux_scale : 0
  int use_drop; DropCfg drop;
  int split_k;                       // > 1: fp32 red.add into C
  int tiles_m, tiles_n, kb_total, kb_per;
};

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// epilogue math on 4 consecutive columns of one row, then the store — all accesses row-contiguous
__device__ __forceinline__ void epilogue_store4(const TcEpilogue& ep, float4 v, int row, int col, const float4& b, uint64_t key,
                                                const float4& old, const uint2& auxpk) {
  v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  if (ep.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  if (ep.use_drop) {      // same stream as drop_mult(): one 64-bit hash per group of 4 elements
    const uint64_t idx = (uint64_t)row * ep.N + col;
    const uint64_t r = mmnas_mix64(key ^ ((idx >> 2) * 0x9E3779B97F4A7C15ull));
    v.x *= ((unsigned)(r) & 0xFFFFu) < ep.drop.thresh ? 0.f : ep.drop.scale;
    v.y *= ((unsigned)(r >> 16) & 0xFFFFu) < ep.drop.thresh ? 0.f : ep.drop.scale;
    v.z *= ((unsigned)(r >> 32) & 0xFFFFu) < ep.drop.thresh ? 0.f : ep.drop.scale;
    v.w *= ((unsigned)(r >> 48) & 0xFFFFu) < ep.drop.thresh ? 0.f : ep.drop.scale;
  }
  if (ep.aux) {
    const uint2 pk = auxpk;
    const __nv_bfloat162 a01 = *reinterpret_cast<const __nv_bfloat162*>(&pk.x);
    const __nv_bfloat162 a23 = *reinterpret_cast<const __nv_bfloat162*>(&pk.y);
    v.x = __low2float(a01) > 0.f ? v.x * ep.aux_scale : 0.f;
    v.y = __high2float(a01) > 0.f ? v.y * ep.aux_scale : 0.f;
    v.z = __low2float(a23) > 0.f ? v.z * ep.aux_scale : 0.f;
    v.w = __high2float(a23) > 0.f ? v.w * ep.aux_scale : 0.f;
  }
  if (ep.out_bf16) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(ep.C) + (long)row * ep.ldc + col) = pk;
  } else {
    float* cp = reinterpret_cast<float*>(ep.C) + (long)row * ep.ldc + col;
    if (ep.split_k > 1) {
      red_add_v4(cp, v);
    } else {
      v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;    // zeros unless ep.accumulate
      *reinterpret_cast<float4*>(cp) = v;
    }
  }
}

template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, TcEpilogue ep) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* staging = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + STG_BYTES);
  // bars: [0,S) full, [S,2S) empty, [2S,2S+2) tmem_full, [2S+2,2S+4) tmem_empty; then the TMEM base word
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = ep.tiles_m * ep.tiles_n * ep.split_k;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // unit -> (split z, m-tile, n-tile), n fastest
  auto decode = [&](int u, int& z, int& m0, int& n0, int& kb0, int& nkb) {
    const int tn = u % ep.tiles_n;
    const int rest = u / ep.tiles_n;
    const int tm = rest % ep.tiles_m;
    z = rest / ep.tiles_m;
    m0 = tm * BM; n0 = tn * BN;
    kb0 = z * ep.kb_per;
    nkb = min(ep.kb_total, kb0 + ep.kb_per) - kb0;
  };

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    pdl_wait();                                   // operands come from the previous kernel(s) of the stream
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int z, m0, n0, kb0, nkb;
      decode(u, z, m0, n0, kb0, nkb);
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % STAGES;
        mbar_wait(empty_bar(s), ((it / STAGES) & 1) ^ 1);
        mbar_expect_tx(full_bar(s), C::STAGE_BYTES);
        const uint32_t sa = smem_base + s * C::STAGE_BYTES, sb = sa + A_TILE_BYTES;
        const int k0 = (kb0 + i) * BK;
        if (!A_MN) {
          tma_load_2d(sa, &tmap_a, full_bar(s), k0, m0);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (BK * 128), &tmap_a, full_bar(s), m0 + 64 * c, k0);
        }
        if (!B_MN) {
          tma_load_2d(sb, &tmap_b, full_bar(s), k0, n0);
        } else {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (BK * 128), &tmap_b, full_bar(s), n0 + 64 * c, k0);
        }
      }
    }
    pdl_launch();                                 // all loads issued: the next kernel may start its prologue
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc(A_MN, B_MN, BM, BN);
    uint32_t it = 0, j = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++j) {
      int z, m0, n0, kb0, nkb;
      decode(u, z, m0, n0, kb0, nkb);
      const uint32_t buf = j & 1;
      mbar_wait(tempty_bar(buf), ((j >> 1) & 1) ^ 1);          // epilogue has drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + buf * BN;
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % STAGES;
        mbar_wait(full_bar(s), (it / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_base + s * C::STAGE_BYTES, sb = sa + A_TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          // K-major: 16 k-elements = 32 B inside the 128 B swizzle row; SBO = 8 rows x 128 B.
          // MN-major: 16 k-rows x 128 B = 2048 B; SBO = 8 k-rows x 128 B, LBO = next 64-wide MN chunk.
          const uint64_t adesc = A_MN ? make_smem_desc(sa + kk * 2048, BK * 128, 1024) : make_smem_desc(sa + kk * 32, 16, 1024);
          const uint64_t bdesc = B_MN ? make_smem_desc(sb + kk * 2048, BK * 128, 1024) : make_smem_desc(sb + kk * 32, 16, 1024);
          umma_bf16(tmem_d, adesc, bdesc, idesc, (i | kk) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(s));        // frees the smem stage when these MMAs retire
      }
      umma_commit(tfull_bar(buf));        // accumulator of this unit complete
    }
  } else if (warp >= 2) {
    // ===== epilogue: TMEM -> registers -> smem staging -> coalesced global =====
    // Eight warps: warp%4 selects the TMEM lane quarter (hardware rule), the two warps of a quarter alternate
    // over the 32-column chunks, so the epilogue keeps up with a 128 x BN x 512 main loop.
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter+32)
    const int half = (warp - 2) >> 2;             // 0: warps 2..5, 1: warps 6..9
    pdl_wait();                                   // this warp reads (accumulate / aux / bias) and writes global memory
    float* stg = staging + (warp - 2) * (32 * STG_PITCH);
    const uint64_t key = ep.use_drop ? drop_key(ep.drop) : 0;
    const int sub_r = lane >> 3, col4 = (lane & 7) * 4;
    uint32_t j = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++j) {
      int z, m0, n0, kb0, nkb;
      decode(u, z, m0, n0, kb0, nkb);
      const uint32_t buf = j & 1;
      mbar_wait(tfull_bar(buf), (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const bool add_bias = ep.bias != nullptr && z == 0;
      const int row_base = m0 + quarter * 32;
#pragma unroll 1
      for (int cc = half; cc < BN / 32; cc += 2) {
        const int col0 = n0 + cc * 32;
        if (col0 >= ep.N) break;                  // warp-uniform (N % 32 == 0)
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + cc * 32), r);
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<uint4*>(stg + lane * STG_PITCH + 4 * g) = make_uint4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
        __syncwarp();
        // all global reads of the chunk (old C for accumulate, the ReLU-mask activation, bias) are issued before
        // any dependent math, so their latencies overlap instead of serialising 8 load->add->store chains
        const int col = col0 + col4;
        float4 bia = make_float4(0.f, 0.f, 0.f, 0.f);
        if (add_bias) bia = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
        float4 old[8];
        uint2 auxpk[8];
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {
          const int row = row_base + itr * 4 + sub_r;
          old[itr] = make_float4(0.f, 0.f, 0.f, 0.f);
          auxpk[itr] = make_uint2(0u, 0u);
          if (row < ep.M) {
            if (ep.accumulate) old[itr] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.C) + (long)row * ep.ldc + col);
            if (ep.aux) auxpk[itr] = __ldg(reinterpret_cast<const uint2*>(ep.aux + (long)row * ep.ld_aux + col));
          }
        }
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {
          const int rr = itr * 4 + sub_r;
          const int row = row_base + rr;
          const float4 v = *reinterpret_cast<const float4*>(stg + rr * STG_PITCH + col4);
          if (row < ep.M) epilogue_store4(ep, v, row, col, bia, key, old[itr], auxpk[itr]);
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

template <bool A_MN, bool B_MN, int BN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const TcEpilogue& ep, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    MMNAS_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_kernel<A_MN, B_MN, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
    attr_done = true;
  }
  const int units = ep.tiles_m * ep.tiles_n * ep.split_k;
  const int grid = units < num_sms() ? units : num_sms();
  MMNAS_CUDA(mmnas_launch(gemm_bf16_tc_kernel<A_MN, B_MN, BN>, dim3(grid), dim3(NUM_THREADS), Cfg<BN>::SMEM_BYTES, s, ta, tb, ep));
  return MMNAS_OK;
}

template <int BN>
int dispatch(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const TcEpilogue& ep, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<false, false, BN>(ta, tb, ep, s);
  if (!a_mn && b_mn) return launch<false, true, BN>(ta, tb, ep, s);
  if (a_mn && !b_mn) return launch<true, false, BN>(ta, tb, ep, s);
  return launch<true, true, BN>(ta, tb, ep, s);
}

}  // namespace

extern "C" int mmnas_gemm_bf16(int M, int N, int K, const void* A, long lda, int a_mn_major, const void* B, long ldb,
                               int b_mn_major, void* C, long ldc, int out_bf16, const float* bias, int relu,
                               int accumulate, const void* aux, long ld_aux, float aux_scale, int split_k,
                               const unsigned long long* rng_state, unsigned long long salt, float p,
                               mmnas_stream stream) {
  MMNAS_CHECK_ARG(M >= 0 && N >= 0 && K >= 1, "gemm_bf16: bad size (K must be >= 1)");
  if (M == 0 || N == 0) return MMNAS_OK;
  MMNAS_CHECK_ARG(A && B && C, "gemm_bf16: null operand");
  MMNAS_CHECK_ARG(N % 32 == 0, "gemm_bf16: N must be a multiple of 32");
  MMNAS_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "gemm_bf16: operand pitches must be multiples of 8 elements (16 B)");
  MMNAS_CHECK_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && ((uintptr_t)C % 16) == 0, "gemm_bf16: 16-byte alignment");
  MMNAS_CHECK_ARG(ldc % (out_bf16 ? 8 : 4) == 0, "gemm_bf16: ldc alignment");
  MMNAS_CHECK_ARG(!bias || ((uintptr_t)bias % 16) == 0, "gemm_bf16: bias alignment");
  if (split_k < 1) split_k = 1;
  const int kb_total = ceil_div(K, BK);
  if (split_k > kb_total) split_k = kb_total;
  MMNAS_CHECK_ARG(split_k == 1 || (!out_bf16 && !relu && !aux && !accumulate && !(p > 0.f)),
                  "gemm_bf16: split-K only supports the plain / bias fp32 epilogue");
  MMNAS_CHECK_ARG(!accumulate || !out_bf16, "gemm_bf16: accumulate needs fp32 output");
  MMNAS_CHECK_ARG(!aux || (ld_aux % 8 == 0 && ((uintptr_t)aux % 8) == 0), "gemm_bf16: aux pitch / alignment");
  int kb_per = ceil_div(kb_total, split_k);
  split_k = ceil_div(kb_total, kb_per);                 // no empty splits
  // BN = 256 halves operand traffic per FLOP; take it unless it costs an extra wave (or multiplies split-K atomics)
  const int tiles_m = ceil_div(M, BM);
  const int sms = num_sms();
  int bn = 128;
  if (split_k == 1 && N % 256 == 0) {
    const int w128 = ceil_div(tiles_m * ceil_div(N, 128), sms);
    const int w256 = 2 * ceil_div(tiles_m * ceil_div(N, 256), sms);
    if (w256 <= w128) bn = 256;
  }
  if (const char* e = getenv("MMNAS_GEMM_BN")) {        // tuning / A-B experiments only
    const int forced = atoi(e);
    if (forced == 128 || (forced == 256 && split_k == 1 && N % 256 == 0)) bn = forced;
  }
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn_major) rc = encode_2d(&ta, A, K, M, lda, BK, BM);     // [M rows][K contiguous]
  else rc = encode_2d(&ta, A, M, K, lda, 64, BK);                 // [K rows][M contiguous]
  if (rc) return rc;
  if (!b_mn_major) rc = encode_2d(&tb, B, K, N, ldb, BK, bn);     // [N rows][K contiguous]
  else rc = encode_2d(&tb, B, N, K, ldb, 64, BK);                 // [K rows][N contiguous]
  if (rc) return rc;
  TcEpilogue ep = {};
  ep.M = M; ep.N = N; ep.K = K; ep.C = C; ep.ldc = ldc; ep.out_bf16 = out_bf16; ep.bias = bias; ep.relu = relu;
  ep.accumulate = accumulate; ep.aux = (const __nv_bfloat16*)aux; ep.ld_aux = ld_aux; ep.aux_scale = aux_scale;
  ep.split_k = split_k;
  ep.use_drop = (p > 0.f && rng_state) ? 1 : 0;
  ep.drop.state = rng_state; ep.drop.salt = salt;
  ep.drop.thresh = (unsigned)(p * 65536.f + 0.5f); ep.drop.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  ep.tiles_m = tiles_m; ep.tiles_n = ceil_div(N, bn); ep.kb_total = kb_total; ep.kb_per = kb_per;
  cudaStream_t s = (cudaStream_t)stream;
  if (bn == 256) return dispatch<256>(a_mn_major, b_mn_major, ta, tb, ep, s);
  return dispatch<128>(a_mn_major, b_mn_major, ta, tb, ep, s);
}
