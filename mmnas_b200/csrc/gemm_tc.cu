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
//   Epilogue: tcgen05.ld (each warp owns the 32 TMEM lanes of its warp%4 quarter; a lane holds one row x 32 columns)
//   -> fused math in registers (+bias, ReLU, dropout with one hash per 4 elements, ReLU-mask of a saved activation)
//   -> 128B-swizzled smem box -> TMA tile store (bf16 / fp32) or TMA f32 reduce-add (accumulate, split-K).
//   BN = 256 (single 128x256x16 UMMA, half the operand traffic per FLOP) when it does not cost a wave.
#include <cstdlib>
#include "tc_common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 320;                      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int EPI_WARPS = 8;
constexpr int STG_BOX_BYTES = 32 * 128;               // one TMA-store box: 32 rows x 128 B (64 bf16 / 32 fp32 columns)
constexpr int STG_BYTES = EPI_WARPS * 2 * STG_BOX_BYTES;    // two boxes per epilogue warp (store i+1 is built while i drains)

template <int BN> struct Cfg {
  static constexpr int STAGES = BN == 64 ? 6 : BN == 128 ? 5 : 3;
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
  int debug;                         // tuning only (MMNAS_GEMM_DEBUG): 2 = no MMA issue, 3 = no epilogue work
};

// ---- epilogue: TMEM -> registers -> (bias, ReLU, dropout, ReLU-mask) -> 128B-swizzled smem box -> TMA store ----------
// tcgen05.ld 32x32b gives every lane ONE accumulator row x 32 consecutive columns, which is already the layout all the
// fused epilogue math wants (the dropout hash covers 4 consecutive columns of a row).  The row is written as 16-byte
// chunks into a [32 rows][128 B] box with the TMA 128B swizzle (chunk ^= row & 7: conflict-free for STS.128), and one
// lane hands the box to the TMA unit: plain tile store for bf16 / fp32 outputs, f32 reduce-add for `accumulate`
// and for split-K.  Row and column clipping at the tensor edge is done by the TMA unit.  The previous epilogue
// (padded smem transpose + per-element address math and predicates) issued ~545 instructions per 32x32 chunk and
// made every projection epilogue-bound at 2x the main-loop time; this one issues ~100.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// hint: bring a tile into L2 (no shared-memory destination, no barrier); legal before griddepcontrol.wait because L2
// is the coherence point — a line the predecessor rewrites afterwards is simply updated in place
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// fused math on this lane's 32 consecutive columns [col0, col0+32) of row `row`
__device__ __forceinline__ void epilogue_math(const TcEpilogue& ep, float (&v)[32], int row, int col0, bool add_bias, uint64_t key,
                                              const uint4 (&aux)[4]) {
  if (add_bias) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + g);     // same address in every lane
      v[4 * g] += b.x; v[4 * g + 1] += b.y; v[4 * g + 2] += b.z; v[4 * g + 3] += b.w;
    }
  }
  if (ep.relu) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (ep.use_drop) {      // same stream as drop_mult(): one 64-bit hash per group of 4 elements
    const uint64_t idx4 = ((uint64_t)row * ep.N + col0) >> 2;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const uint64_t r = mmnas_mix64(key ^ ((idx4 + g) * 0x9E3779B97F4A7C15ull));
      v[4 * g] *= ((unsigned)(r) & 0xFFFFu) < ep.drop.thresh ? 0.f : ep.drop.scale;
      v[4 * g + 1] *= ((unsigned)(r >> 16) & 0xFFFFu) < ep.drop.thresh ? 0.f : ep.drop.scale;
      v[4 * g + 2] *= ((unsigned)(r >> 32) & 0xFFFFu) < ep.drop.thresh ? 0.f : ep.drop.scale;
      v[4 * g + 3] *= ((unsigned)(r >> 48) & 0xFFFFu) < ep.drop.thresh ? 0.f : ep.drop.scale;
    }
  }
  if (ep.aux) {           // v = aux > 0 ? v * aux_scale : 0  (backward through ReLU + dropout-scale of the FFN hidden layer)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t w[4] = {aux[q].x, aux[q].y, aux[q].z, aux[q].w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const __nv_bfloat162 a2 = *reinterpret_cast<const __nv_bfloat162*>(&w[h]);
        const int i = 8 * q + 2 * h;
        v[i] = __low2float(a2) > 0.f ? v[i] * ep.aux_scale : 0.f;
        v[i + 1] = __high2float(a2) > 0.f ? v[i + 1] * ep.aux_scale : 0.f;
      }
    }
  }
}

// One accumulator tile of 32 rows (this warp's TMEM lane quarter) x BN columns.  `tmem_acc` = TMEM address of column 0
// of the accumulator in this warp's lane quarter; the two warps of a quarter (half = 0 / 1) alternate over the store
// boxes (64 columns for bf16, 32 for fp32).  `stg` = shared address of this warp's two boxes, `sel` = which is next.
template <int BN>
__device__ __forceinline__ void epilogue_unit(const TcEpilogue& ep, const CUtensorMap* tmap_c, uint32_t tmem_acc, uint32_t stg,
                                              uint32_t& sel, int row0, int n0, int half, bool add_bias, uint64_t key, int lane) {
  const int row = row0 + lane;
  const bool row_ok = row < ep.M;
  const bool reduce = ep.accumulate || ep.split_k > 1;
  const uint32_t swz = (uint32_t)(lane & 7);
  const int chunks_per_box = ep.out_bf16 ? 2 : 1;
  const int boxes = BN / 32 / chunks_per_box;
#pragma unroll 1
  for (int bx = half; bx < boxes; bx += 2) {
    const int colb = n0 + bx * chunks_per_box * 32;
    if (colb >= ep.N) break;                  // warp-uniform
    const uint32_t box = stg + sel * STG_BOX_BYTES;
    const uint32_t rowaddr = box + (uint32_t)lane * 128u;
    if (lane == 0) bulk_wait_read<1>();       // the store issued two boxes ago has finished reading this buffer
    __syncwarp();
#pragma unroll 1
    for (int c = 0; c < chunks_per_box; ++c) {
      const int col0 = colb + c * 32;
      if (col0 >= ep.N) break;                // N % 64 == 32: the TMA unit clips the missing half of the box
      uint4 aux[4] = {};
      if (ep.aux && row_ok) {
        const uint4* ap = reinterpret_cast<const uint4*>(ep.aux + (long)row * ep.ld_aux + col0);
#pragma unroll
        for (int q = 0; q < 4; ++q) aux[q] = __ldg(ap + q);
      }
      uint32_t r[32];
      tmem_ld32(tmem_acc + (uint32_t)((bx * chunks_per_box + c) * 32), r);
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      epilogue_math(ep, v, row, col0, add_bias, key, aux);
      if (ep.out_bf16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {         // 16-byte chunk (4c + q) of the 128-byte row: 8 bf16 columns
          uint32_t pk[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const __nv_bfloat162 t = __floats2bfloat162_rn(v[8 * q + 2 * h], v[8 * q + 2 * h + 1]);
            pk[h] = *reinterpret_cast<const uint32_t*>(&t);
          }
          sts128(rowaddr + ((((uint32_t)(4 * c + q)) ^ swz) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q)           // 16-byte chunk q: 4 fp32 columns
          sts128(rowaddr + (((uint32_t)q ^ swz) << 4), __float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]),
                 __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA unit
    __syncwarp();
    if (lane == 0) {
      if (reduce) tma_reduce_add_2d(tmap_c, box, colb, row0);
      else tma_store_2d(tmap_c, box, colb, row0);
      bulk_commit();
    }
    sel ^= 1u;
  }
}

template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, TcEpilogue ep) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by POINTER arithmetic on the shared array: an integer round trip loses the address space
  // and every staging access becomes a generic ST.E / LD.E instead of STS / LDS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t staging = smem_u32(smem) + STAGES * C::STAGE_BYTES;      // 1024-byte aligned (swizzle atom)
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
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
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
  // Let the next kernel of the stream be scheduled now: its CTAs land on SMs this grid leaves idle (28..112 CTAs for
  // the text-stream projections) or free first, run their own prologue and park in griddepcontrol.wait.
  if (threadIdx.x == 0) pdl_launch();

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
    if (blockIdx.x < units) {                     // warm L2 with the first stages of B (weights: cold, a step old)
      int z, m0, n0, kb0, nkb;
      decode(blockIdx.x, z, m0, n0, kb0, nkb);
      const int pre = nkb < STAGES ? nkb : STAGES;
      for (int i = 0; i < pre; ++i) {
        const int k0 = (kb0 + i) * BK;
        if (!B_MN) {
          tma_prefetch_2d(&tmap_b, k0, n0);
        } else {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_prefetch_2d(&tmap_b, n0 + 64 * c, k0);
        }
      }
    }
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
          if (ep.debug != 2) umma_bf16(tmem_d, adesc, bdesc, idesc, (i | kk) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(s));        // frees the smem stage when these MMAs retire
      }
      umma_commit(tfull_bar(buf));        // accumulator of this unit complete
    }
  } else if (warp >= 2) {
    // ===== epilogue =====
    // Eight warps: warp%4 selects the TMEM lane quarter (hardware rule), the two warps of a quarter alternate
    // over the store boxes, so the epilogue keeps up with a 128 x BN x 512 main loop.
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter+32)
    const int half = (warp - 2) >> 2;             // 0: warps 2..5, 1: warps 6..9
    pdl_wait();                                   // this warp reads (aux / bias) and writes global memory
    const uint32_t stg = staging + (uint32_t)(warp - 2) * (2 * STG_BOX_BYTES);
    const uint64_t key = ep.use_drop ? drop_key(ep.drop) : 0;
    uint32_t j = 0, sel = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++j) {
      int z, m0, n0, kb0, nkb;
      decode(u, z, m0, n0, kb0, nkb);
      const uint32_t buf = j & 1;
      mbar_wait(tfull_bar(buf), (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (ep.debug != 3)
        epilogue_unit<BN>(ep, &tmap_c, tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN), stg, sel,
                          m0 + quarter * 32, n0, half, ep.bias != nullptr && z == 0, key, lane);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
    if (lane == 0) bulk_wait_read<0>();           // the TMA unit has read the last boxes: shared memory may be released
                                                  // (the writes themselves complete before the grid does)
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs on one TPC owns a 256 x 256 output tile.  Each CTA
// stages its own 128 rows of A and its own 128 of the 256 B rows; ONE 256x256x16 MMA issued by the leader reads
// both CTAs' shared memory and writes 128 accumulator rows into each CTA's TMEM.  Per 128x256 of output a CTA pulls
// 32 KB per k-block through L2 instead of 48 KB (1-CTA, BN=256) or 64 KB (BN=128) — the projections of this model
// (K = 512) are bound by exactly that L2->SM traffic, not by the tensor pipe (GEMM section).
//   full[s]   : leader's barrier; the leader arms it with the bytes of BOTH CTAs, both CTAs' TMA complete on it
//   empty[s]  : one per CTA, released by the leader's multicast commit once the MMAs reading the stage retired
//   tfull[b]  : one per CTA, multicast commit after the last MMA of a unit
//   tempty[b] : leader's barrier, 2 x EPI_WARPS arrivals (the peer's epilogue warps arrive remotely)
// ---------------------------------------------------------------------------------------------------------
constexpr int PAIR_BN = 256;                          // N of the pair tile; each CTA stages PAIR_BN / 2 rows of B
constexpr int PAIR_STAGES = 5;
constexpr int PAIR_STAGE_BYTES = A_TILE_BYTES + (PAIR_BN / 2) * BK * 2;
constexpr int PAIR_SMEM_BYTES = PAIR_STAGES * PAIR_STAGE_BYTES + STG_BYTES + 1024 + 256;
constexpr uint32_t PAIR_TMEM_COLS = 2 * PAIR_BN;

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_c, TcEpilogue ep) {
  constexpr int STAGES = PAIR_STAGES;
  extern __shared__ uint8_t smem_raw[];
  // both CTAs must use identical shared-memory offsets: the MMA applies the leader's descriptors in the peer too
  // 1024-byte alignment by POINTER arithmetic on the shared array: an integer round trip loses the address space
  // and every staging access becomes a generic ST.E / LD.E instead of STS / LDS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t staging = smem_u32(smem) + STAGES * PAIR_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * PAIR_STAGE_BYTES + STG_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int units = ep.tiles_m * ep.tiles_n * ep.split_k;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 2 * EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {      // one warp of EACH CTA: the pair allocation is collective
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(PAIR_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                               // the peer's barriers exist before anything signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_launch();

  auto decode = [&](int u, int& z, int& m0, int& n0, int& kb0, int& nkb) {
    const int tn = u % ep.tiles_n;
    const int rest = u / ep.tiles_n;
    const int tm = rest % ep.tiles_m;
    z = rest / ep.tiles_m;
    m0 = tm * (2 * BM) + (int)rank * BM;            // this CTA's 128 rows of the 256-row tile
    n0 = tn * PAIR_BN;
    kb0 = z * ep.kb_per;
    nkb = min(ep.kb_total, kb0 + ep.kb_per) - kb0;
  };

  if (warp == 0 && lane == 0) {
    // ===== TMA producer (both CTAs) =====
    pdl_wait();
    uint32_t it = 0;
    for (int u = cluster_id; u < units; u += n_clusters) {
      int z, m0, n0, kb0, nkb;
      decode(u, z, m0, n0, kb0, nkb);
      const int nh = n0 + (int)rank * (PAIR_BN / 2);           // this CTA's half of the B rows
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % STAGES;
        mbar_wait(empty_bar(s), ((it / STAGES) & 1) ^ 1);
        if (leader) mbar_expect_tx(full_bar(s), 2 * PAIR_STAGE_BYTES);
        const uint32_t fb = mapa_shared(full_bar(s), 0);
        const uint32_t sa = smem_base + s * PAIR_STAGE_BYTES, sb = sa + A_TILE_BYTES;
        const int k0 = (kb0 + i) * BK;
        if (!A_MN) {
          tma_load_2d_pair(sa, &tmap_a, fb, k0, m0);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c) tma_load_2d_pair(sa + c * (BK * 128), &tmap_a, fb, m0 + 64 * c, k0);
        }
        if (!B_MN) {
          tma_load_2d_pair(sb, &tmap_b, fb, k0, nh);
        } else {
#pragma unroll
          for (int c = 0; c < PAIR_BN / 128; ++c) tma_load_2d_pair(sb + c * (BK * 128), &tmap_b, fb, nh + 64 * c, k0);
        }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===== MMA issuer (leader CTA only) =====
    constexpr uint32_t idesc = make_idesc(A_MN, B_MN, 2 * BM, PAIR_BN);
    uint32_t it = 0, j = 0;
    for (int u = cluster_id; u < units; u += n_clusters, ++j) {
      int z, m0, n0, kb0, nkb;
      decode(u, z, m0, n0, kb0, nkb);
      const uint32_t buf = j & 1;
      mbar_wait_cluster(tempty_bar(buf), ((j >> 1) & 1) ^ 1);  // both CTAs' epilogues drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + buf * PAIR_BN;
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % STAGES;
        mbar_wait(full_bar(s), (it / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_base + s * PAIR_STAGE_BYTES, sb = sa + A_TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t adesc = A_MN ? make_smem_desc(sa + kk * 2048, BK * 128, 1024) : make_smem_desc(sa + kk * 32, 16, 1024);
          const uint64_t bdesc = B_MN ? make_smem_desc(sb + kk * 2048, BK * 128, 1024) : make_smem_desc(sb + kk * 32, 16, 1024);
          if (ep.debug != 2) umma_bf16_pair(tmem_d, adesc, bdesc, idesc, (i | kk) != 0 ? 1u : 0u);
        }
        umma_commit_pair(empty_bar(s), 3);
      }
      umma_commit_pair(tfull_bar(buf), 3);
    }
  } else if (warp >= 2) {
    // ===== epilogue (both CTAs, each its own 128 accumulator rows) =====
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    pdl_wait();
    const uint32_t stg = staging + (uint32_t)(warp - 2) * (2 * STG_BOX_BYTES);
    const uint64_t key = ep.use_drop ? drop_key(ep.drop) : 0;
    const uint32_t te0 = mapa_shared(tempty_bar(0), 0), te1 = mapa_shared(tempty_bar(1), 0);
    uint32_t j = 0, sel = 0;
    for (int u = cluster_id; u < units; u += n_clusters, ++j) {
      int z, m0, n0, kb0, nkb;
      decode(u, z, m0, n0, kb0, nkb);
      const uint32_t buf = j & 1;
      mbar_wait(tfull_bar(buf), (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (ep.debug != 3)
        epilogue_unit<PAIR_BN>(ep, &tmap_c, tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * PAIR_BN), stg, sel,
                               m0 + quarter * 32, n0, half, ep.bias != nullptr && z == 0, key, lane);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) mbar_arrive_cluster(buf ? te1 : te0);
    }
    if (lane == 0) bulk_wait_read<0>();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                               // no CTA leaves (or frees TMEM) while its peer still works
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(PAIR_TMEM_COLS) : "memory");
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
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const TcEpilogue& ep, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    MMNAS_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_kernel<A_MN, B_MN, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
    attr_done = true;
  }
  const int units = ep.tiles_m * ep.tiles_n * ep.split_k;
  const int grid = units < num_sms() ? units : num_sms();
  MMNAS_CUDA(mmnas_launch(gemm_bf16_tc_kernel<A_MN, B_MN, BN>, dim3(grid), dim3(NUM_THREADS), Cfg<BN>::SMEM_BYTES, s, ta, tb, tc, ep));
  return MMNAS_OK;
}

template <int BN>
int dispatch(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const TcEpilogue& ep, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<false, false, BN>(ta, tb, tc, ep, s);
  if (!a_mn && b_mn) return launch<false, true, BN>(ta, tb, tc, ep, s);
  if (a_mn && !b_mn) return launch<true, false, BN>(ta, tb, tc, ep, s);
  return launch<true, true, BN>(ta, tb, tc, ep, s);
}

template <bool A_MN, bool B_MN>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const TcEpilogue& ep, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    MMNAS_CUDA(cudaFuncSetAttribute(gemm_bf16_pair_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES));
    attr_done = true;
  }
  const int units = ep.tiles_m * ep.tiles_n * ep.split_k;
  const int clusters = units < num_sms() / 2 ? units : num_sms() / 2;
  MMNAS_CUDA(mmnas_launch(gemm_bf16_pair_kernel<A_MN, B_MN>, dim3(2 * clusters), dim3(NUM_THREADS), PAIR_SMEM_BYTES, s, ta, tb, tc, ep));
  return MMNAS_OK;
}

int dispatch_pair(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const TcEpilogue& ep, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_pair<false, false>(ta, tb, tc, ep, s);
  if (!a_mn && b_mn) return launch_pair<false, true>(ta, tb, tc, ep, s);
  if (a_mn && !b_mn) return launch_pair<true, false>(ta, tb, tc, ep, s);
  return launch_pair<true, true>(ta, tb, tc, ep, s);
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
  // Tile configuration from a small cost model fitted on B200 (scripts/bench_gemm_pair.py, K = 512 projections of the
  // step): time = fixed + rounds * per-round cost * K/512, rounds = ceil(work units / resident CTAs or clusters).
  //   single CTA, 128x128 tiles : 5.5 us + 2.7 us per round      single CTA, 128x256 : 5.5 us + 3.6 us per round
  //   CTA pair, 256x256 tiles   : 7.3 us + 2.8 us per round (cluster launch + two cluster syncs cost ~1.8 us)
  // Split-K launches (weight gradients) take the pair kernel once they fill >= 48 of the 74 clusters.
  const int tiles_m = ceil_div(M, BM);
  const int sms = num_sms();
  const bool n256 = (N % 256 == 0);
  int bn = 128;
  bool pair = false;
  if (split_k == 1) {
    const float kf = (float)K / 512.f;
    const float t128 = 5.5f + 2.7f * kf * ceil_div(tiles_m * ceil_div(N, 128), sms);
    const float t256 = n256 ? 5.5f + 3.6f * kf * ceil_div(tiles_m * (N / 256), sms) : 1e30f;
    const float tpair = (n256 && M >= 2 * BM) ? 7.3f + 2.8f * kf * ceil_div(ceil_div(M, 2 * BM) * (N / 256), sms / 2) : 1e30f;
    if (t256 < t128) bn = 256;
    pair = tpair < (t256 < t128 ? t256 : t128);
    // Few row blocks (the 14-token text stream: 7 of them) and at most half a wave of 128-wide tiles: 64-wide tiles
    // double the CTAs at 24 KB instead of 32 KB per CTA per k-block (6.4 -> 6.0 us on 896 x 512 x 512).
    if (!pair && bn == 128 && N % 64 == 0 && tiles_m * ceil_div(N, 128) * 2 <= sms) bn = 64;
  } else {
    pair = n256 && M >= 2 * BM && ceil_div(M, 2 * BM) * (N / PAIR_BN) * split_k >= 48;
  }
  if (const char* e = getenv("MMNAS_GEMM_BN")) {        // tuning / A-B experiments only
    const int forced = atoi(e);
    if (forced == 128 || (forced == 256 && split_k == 1 && n256) || (forced == 64 && N % 64 == 0)) bn = forced;
  }
  if (const char* e = getenv("MMNAS_GEMM_PAIR")) {      // tuning / A-B experiments only: 0 = never, 1 = whenever legal
    pair = atoi(e) != 0 && n256 && M >= 2 * BM;
  }
  if (pair) bn = PAIR_BN / 2;                           // rows of B each CTA stages
  CUtensorMap ta, tb, tc;
  int rc;
  // output boxes of 32 rows x 128 bytes: 64 bf16 or 32 fp32 columns
  rc = out_bf16 ? encode_2d(&tc, C, N, M, ldc, 64, 32) : encode_2d_f32(&tc, C, N, M, ldc, 32, 32);
  if (rc) return rc;
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
  if (pair) { ep.tiles_m = ceil_div(M, 2 * BM); ep.tiles_n = N / PAIR_BN; }
  if (const char* e = getenv("MMNAS_GEMM_DEBUG")) ep.debug = atoi(e);
  cudaStream_t s = (cudaStream_t)stream;
  if (pair) return dispatch_pair(a_mn_major, b_mn_major, ta, tb, tc, ep, s);
  if (bn == 256) return dispatch<256>(a_mn_major, b_mn_major, ta, tb, tc, ep, s);
  if (bn == 64) return dispatch<64>(a_mn_major, b_mn_major, ta, tb, tc, ep, s);
  return dispatch<128>(a_mn_major, b_mn_major, ta, tb, tc, ep, s);
}
