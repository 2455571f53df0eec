// Projection GEMM with the block tail fused into its epilogue (bf16 arm):
//     z   = x + dropout(A W^T + bias)                    residual add of the (dropped-out) branch   modules.py:261-271
//     out = a_2 * (z - mean) / (std_unbiased + eps) + b_2                                          modules.py:52-56
// i.e. the merge projection of SelfAtt / GuidedAtt / RelSelfAtt (modules.py:199,240) or the second FFN layer
// (modules.py:38,41) together with the residual + LayerNorm kernel that used to follow it: the fp32 branch [M,H] no
// longer makes a round trip through HBM and one launch per block disappears.
//
// LayerNorm needs whole rows, and a whole H = 512 row of fp32 accumulators is all 512 TMEM columns of one SM.  So a
// row block is owned by a CLUSTER: CTA pairs (tcgen05 cta_group::2, one 256 x 256 x 16 MMA per issue, each CTA holding
// 128 rows x 256 accumulator columns and staging only half of the B tile) side by side along N —
//     H = 256: cluster of 2 (one pair),   H = 512: cluster of 4 (two pairs),
// and the row statistics are completed across the pairs through distributed shared memory.  The statistics are the
// exact two-pass ones of the stand-alone kernel (mean, then the centred sum of squares): z is written back into TMEM
// (tcgen05.st) after the first pass and re-read for the second and third — TMEM bandwidth makes the extra passes free.
//   pass 1  acc (+bias) -> dropout -> + x -> z : tcgen05.st, TMA tile store of z (saved for the backward), row sums
//   pass 2  centred sums of squares -> sigma
//   pass 3  normalise, scale / shift -> TMA tile stores of out (fp32) and of its bf16 copy (next block's operand)
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (leader CTA of each pair), warps 2..17 epilogue (thread =
// accumulator row; the four warps of a TMEM lane quarter split the 256 columns — the epilogue is instruction-bound, so it
// gets every warp the register file allows).  One row block per cluster:
// M = 6400 gives 25 clusters of 4 = 100 CTAs, one wave.  The stage ring is reused as TMA-store staging once the main
// loop has drained.
#include <cstdlib>
#include "tc_common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int BM = 128, BK = 64, PN = 256;             // rows per CTA, k-block, columns per pair (= per CTA accumulator)
constexpr int STAGES = 5;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = (PN / 2) * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KB
constexpr int EPI_WARPS = 16;                          // four per TMEM lane quarter: 64 of the 256 columns each
constexpr int NUM_THREADS = (2 + EPI_WARPS) * 32;      // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr int BOX_BYTES = 32 * 128;                    // one TMA-store box: 32 rows x 128 B
constexpr int BOXES_PER_WARP = 3;                      // two fp32 boxes (32 columns each) + one bf16 box (64 columns)
constexpr int STAGING_BYTES = EPI_WARPS * BOXES_PER_WARP * BOX_BYTES;              // 192 KB: the stage ring + an extension
constexpr int RING_BYTES = STAGES * STAGE_BYTES;
constexpr int EXTRA_BYTES = STAGING_BYTES > RING_BYTES ? STAGING_BYTES - RING_BYTES : 0;
constexpr int PART_FLOATS = 6 * BM;                    // row partials: 4 local column quarters, remote sum, remote sum of squares
constexpr int SMEM_BYTES = RING_BYTES + EXTRA_BYTES + 1024 /*align*/ + PART_FLOATS * 4 + 256 /*barriers*/;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct LnArgs {
  int M, N, K, kb_total;
  const float* bias;           // [N] or null
  const float* x;              // residual input [M,N] fp32 or null
  const float *gamma, *beta;   // [N]
  float eps;
  float *mean, *sigma;         // [M]
  int use_drop; DropCfg drop;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }     // the 16 epilogue warps
__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}

// one 32 x 128-byte box of fp32: this lane's 32 columns as eight 16-byte chunks, TMA 128B swizzle
__device__ __forceinline__ void stage_f32_row(uint32_t box, int lane, const float (&v)[32]) {
  const uint32_t rowaddr = box + (uint32_t)lane * 128u, swz = (uint32_t)(lane & 7);
#pragma unroll
  for (int q = 0; q < 8; ++q)
    sts128(rowaddr + (((uint32_t)q ^ swz) << 4), __float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]),
           __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
}
// half of a bf16 box (64 columns per 128-byte row): chunk half c (0 / 1) = this lane's 32 columns
__device__ __forceinline__ void stage_bf16_half(uint32_t box, int lane, int c, const float (&v)[32]) {
  const uint32_t rowaddr = box + (uint32_t)lane * 128u, swz = (uint32_t)(lane & 7);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t pk[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const __nv_bfloat162 t = __floats2bfloat162_rn(v[8 * q + 2 * h], v[8 * q + 2 * h + 1]);
      pk[h] = *reinterpret_cast<const uint32_t*>(&t);
    }
    sts128(rowaddr + ((((uint32_t)(4 * c + q)) ^ swz) << 4), pk[0], pk[1], pk[2], pk[3]);
  }
}

// CN = column pairs per row block (N = 256 CN); cluster = 2 CN CTAs, rank r: pair r >> 1, row half r & 1
template <int CN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_out,
               const __grid_constant__ CUtensorMap tmap_out16, const __grid_constant__ CUtensorMap tmap_x, LnArgs ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  float* part = reinterpret_cast<float*>(smem + RING_BYTES + EXTRA_BYTES);   // [4][BM] local column quarters, [BM] remote x2
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING_BYTES + EXTRA_BYTES + PART_FLOATS * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1 + EPI_WARPS);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
  auto x_bar = [&](int w) { return bar_base + 8u * (2 * STAGES + 1 + w); };      // one per epilogue warp: its residual boxes

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t cp = rank >> 1, pr = rank & 1u;          // column pair, row half inside the pair
  const bool leader = pr == 0;
  const uint32_t lead_rank = rank & ~1u;
  const int tile_m = blockIdx.x / (2 * CN);
  const int m0 = tile_m * (2 * BM) + (int)pr * BM;       // this CTA's 128 rows
  const int n0 = (int)cp * PN;                           // this pair's 256 columns

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_z) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out16) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(x_bar(w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {      // one warp of EACH CTA of the pair: the pair allocation is collective
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)PN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                   // every CTA's barriers exist before anything signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) pdl_launch();

  if (warp == 0 && lane == 0) {
    // ===== TMA producer (every CTA: its 128 rows of A, its 128 of the pair's 256 rows of B) =====
    pdl_wait();
    const int nh = n0 + (int)pr * (PN / 2);
    for (int i = 0; i < ep.kb_total; ++i) {
      const int s = i % STAGES;
      mbar_wait(empty_bar(s), ((i / STAGES) & 1) ^ 1);
      if (leader) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES);
      const uint32_t fb = mapa_shared(full_bar(s), lead_rank);
      const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
      tma_load_2d_pair(sa, &tmap_a, fb, i * BK, m0);
      tma_load_2d_pair(sb, &tmap_b, fb, i * BK, nh);
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===== MMA issuer (leader CTA of the pair) =====
    constexpr uint32_t idesc = make_idesc(false, false, 2 * BM, PN);
    const uint16_t mask = (uint16_t)(3u << lead_rank);
    for (int i = 0; i < ep.kb_total; ++i) {
      const int s = i % STAGES;
      mbar_wait(full_bar(s), (i / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
      for (int kk = 0; kk < BK / 16; ++kk)
        umma_bf16_pair(tmem_base, make_smem_desc(sa + kk * 32, 16, 1024), make_smem_desc(sb + kk * 32, 16, 1024), idesc,
                       (i | kk) != 0 ? 1u : 0u);
      umma_commit_pair(empty_bar(s), mask);
    }
    umma_commit_pair(tfull_bar, mask);
  }

  // ===== epilogue (warps 2..17 of every CTA): thread = accumulator row, 64 columns = two 32-column chunks =====
  const bool epi = warp >= 2;
  const int quarter = warp & 3, sub = epi ? (warp - 2) >> 2 : 0;
  const int rloc = quarter * 32 + lane;                  // row inside the CTA = TMEM lane
  const int row = m0 + rloc;
  const bool row_ok = row < ep.M;
  const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(sub * 64);
  const int colw = n0 + sub * 64;                        // first global column of this warp's 64 columns
  const uint32_t boxes = smem_base + (uint32_t)((warp - 2) * BOXES_PER_WARP) * BOX_BYTES;      // stage ring (+ extension)
  float s1 = 0.f, stat_s = 0.f, stat_q = 0.f;
  if (epi) {
    // The residual tile of this warp (32 rows x 64 columns) arrives by TMA as two 128B-swizzled boxes in the slots that
    // become the z boxes: a thread-per-row global read would touch 32 cache lines per instruction (measured: 9 us of
    // L1 tag traffic per CTA), the TMA unit streams whole rows.  Issued as soon as the stage ring is free.
    // The dropout decisions do not depend on the accumulators: they are hashed while the main loop runs (the epilogue
    // is otherwise bound by exactly these ~40 integer instructions per 4 elements) and kept as one bit per element.
    uint32_t keep[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
    if (ep.use_drop) {        // the stream of ln_fwd_kernel's drop4_mult: one hash per 4 consecutive elements of [M,N]
      const uint64_t key = drop_key(ep.drop);           // {seed, step}: written at the start of the step
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint64_t idx4 = ((uint64_t)row * ep.N + colw + c * 32) >> 2;
        uint32_t m = 0;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint64_t h = mmnas_mix64(key ^ ((idx4 + g) * 0x9E3779B97F4A7C15ull));
#pragma unroll
          for (int u = 0; u < 4; ++u) m |= (((unsigned)(h >> (16 * u)) & 0xFFFFu) < ep.drop.thresh ? 0u : 1u) << (4 * g + u);
        }
        keep[c] = m;
      }
    }
    pdl_wait();
    mbar_wait(tfull_bar, 0);                             // all MMAs of this row block retired: accumulators complete,
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");     // and every stage of the ring is free
    const uint32_t xb = x_bar(warp - 2);
    if (ep.x && lane == 0) {
      mbar_expect_tx(xb, 2 * BOX_BYTES);
      tma_load_2d(boxes, &tmap_x, xb, colw, m0 + quarter * 32);
      tma_load_2d(boxes + BOX_BYTES, &tmap_x, xb, colw + 32, m0 + quarter * 32);
    }
    // ---- pass 1: z = x + dropout(acc + bias); z -> TMEM and -> global; row sums
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int col0 = colw + c * 32;
      uint32_t r[32];
      tmem_ld32_nowait(trow + (uint32_t)(c * 32), r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      if (ep.bias) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + g);
          v[4 * g] += b.x; v[4 * g + 1] += b.y; v[4 * g + 2] += b.z; v[4 * g + 3] += b.w;
        }
      }
      if (ep.use_drop) {
        const uint32_t m = keep[c];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = ((m >> i) & 1u) ? v[i] * ep.drop.scale : 0.f;
      }
      const uint32_t box = boxes + (uint32_t)c * BOX_BYTES;
      if (ep.x) {
        if (c == 0) mbar_wait(xb, 0);                    // both residual boxes have landed
        const uint32_t rowaddr = box + (uint32_t)lane * 128u, swz = (uint32_t)(lane & 7);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float4 t;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                       : "r"(rowaddr + (((uint32_t)g ^ swz) << 4)));
          v[4 * g] += t.x; v[4 * g + 1] += t.y; v[4 * g + 2] += t.z; v[4 * g + 3] += t.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) { s1 += v[i]; r[i] = __float_as_uint(v[i]); }
      tmem_st32(trow + (uint32_t)(c * 32), r);
      stage_f32_row(box, lane, v);                       // z overwrites this thread's own residual row
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { tma_store_2d(&tmap_z, box, col0, m0 + quarter * 32); bulk_commit(); }
    }
    tmem_st_wait();
    part[sub * BM + rloc] = s1;
    epi_bar_sync();
    s1 = (part[rloc] + part[BM + rloc]) + (part[2 * BM + rloc] + part[3 * BM + rloc]);      // this CTA's 256 columns
    const float mean_l = s1 * (1.f / (float)PN);
    // ---- pass 2: sum of squares centred on the mean of this CTA's columns (two-pass, like the stand-alone kernel)
    float q2 = 0.f;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(trow + (uint32_t)(c * 32), r);
#pragma unroll
      for (int i = 0; i < 32; ++i) { const float d = __uint_as_float(r[i]) - mean_l; q2 = fmaf(d, d, q2); }
    }
    epi_bar_sync();                                      // everyone has consumed the pass-1 partials
    part[sub * BM + rloc] = q2;
    epi_bar_sync();
    q2 = (part[rloc] + part[BM + rloc]) + (part[2 * BM + rloc] + part[3 * BM + rloc]);
    if (CN > 1 && sub == 0) {                            // ONE exchange with the other pair: (sum, centred sum of squares)
      st_cluster_f32(mapa_shared(smem_u32(part + 4 * BM + rloc), rank ^ 2u), s1);
      st_cluster_f32(mapa_shared(smem_u32(part + 5 * BM + rloc), rank ^ 2u), q2);
    }
    stat_s = s1; stat_q = q2;
  }
  __syncwarp();                                          // lanes that took a single-lane role rejoin their warp
  if (CN > 1) cluster_sync_all();                        // the other pair's statistics have landed (all threads take part)
  if (epi) {
    float mean = stat_s * (1.f / (float)PN), q2 = stat_q;
    if (CN > 1) {   // Chan et al.: M2 = M2_a + M2_b + delta^2 n_a n_b / (n_a + n_b), n_a = n_b = 256
      const float s_r = part[4 * BM + rloc], q_r = part[5 * BM + rloc];
      const float delta = (s_r - stat_s) * (1.f / (float)PN);
      q2 = (q2 + q_r) + delta * delta * (0.5f * (float)PN);
      mean = (stat_s + s_r) * (1.f / (float)(2 * PN));
    }
    const float sigma = sqrtf(q2 / (float)(ep.N - 1));
    const float t = 1.f / (sigma + ep.eps);
    if (sub == 0 && cp == 0 && row_ok) { ep.mean[row] = mean; ep.sigma[row] = sigma; }
    // ---- pass 3: out = gamma (z - mean) t + beta -> two fp32 boxes (slots 0 / 1) + one bf16 box of 64 columns (slot 2)
    if (lane == 0) bulk_wait_read<0>();                  // the z stores of pass 1 have read slots 0 / 1
    __syncwarp();
    const uint32_t box16 = boxes + 2u * BOX_BYTES;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int col0 = colw + c * 32;
      uint32_t r[32];
      tmem_ld32(trow + (uint32_t)(c * 32), r);
      float v[32];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 ga = __ldg(reinterpret_cast<const float4*>(ep.gamma + col0) + g);
        const float4 be = __ldg(reinterpret_cast<const float4*>(ep.beta + col0) + g);
        v[4 * g] = ga.x * (__uint_as_float(r[4 * g]) - mean) * t + be.x;
        v[4 * g + 1] = ga.y * (__uint_as_float(r[4 * g + 1]) - mean) * t + be.y;
        v[4 * g + 2] = ga.z * (__uint_as_float(r[4 * g + 2]) - mean) * t + be.z;
        v[4 * g + 3] = ga.w * (__uint_as_float(r[4 * g + 3]) - mean) * t + be.w;
      }
      const uint32_t box = boxes + (uint32_t)c * BOX_BYTES;
      stage_f32_row(box, lane, v);
      stage_bf16_half(box16, lane, c, v);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { tma_store_2d(&tmap_out, box, col0, m0 + quarter * 32); bulk_commit(); }
    }
    if (lane == 0) { tma_store_2d(&tmap_out16, box16, colw, m0 + quarter * 32); bulk_commit(); bulk_wait_read<0>(); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                    // no CTA leaves (or frees TMEM) while a peer still works
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)PN) : "memory");
  }
}

template <int CN>
int launch_ln(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tz, const CUtensorMap& to, const CUtensorMap& to16,
              const CUtensorMap& tx, const LnArgs& ep, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    MMNAS_CUDA(cudaFuncSetAttribute(gemm_ln_kernel<CN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done = true;
  }
  const int row_blocks = ceil_div(ep.M, 2 * BM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(row_blocks * 2 * CN); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * CN; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mmnas_pdl_enabled() ? 2 : 1;
  mmnas_count_launch();
  MMNAS_CUDA(cudaLaunchKernelEx(&cfg, gemm_ln_kernel<CN>, ta, tb, tz, to, to16, tx, ep));
  return MMNAS_OK;
}

}  // namespace

// Returns MMNAS_ERR_UNSUPPORTED for shapes outside the kernel (the caller then runs GEMM + residual/LayerNorm separately).
extern "C" int mmnas_gemm_ln_bf16(int M, int N, int K, const void* A, long lda, const void* W, long ldb, const float* bias,
                                  const float* x, const float* gamma, const float* beta, float eps, float* z, float* out,
                                  void* out_bf16, float* mean, float* sigma, const unsigned long long* rng_state,
                                  unsigned long long salt, float p, mmnas_stream stream) {
  MMNAS_CHECK_ARG(M >= 0 && K >= 1, "gemm_ln: bad size");
  if (M == 0) return MMNAS_OK;
  if (!(N == 256 || N == 512) || !gamma || !beta || !out_bf16) return MMNAS_ERR_UNSUPPORTED;
  MMNAS_CHECK_ARG(A && W && z && out && mean && sigma, "gemm_ln: null buffer");
  MMNAS_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "gemm_ln: operand alignment");
  MMNAS_CHECK_ARG(((uintptr_t)z % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)out_bf16 % 16) == 0, "gemm_ln: output alignment");
  MMNAS_CHECK_ARG((!bias || ((uintptr_t)bias % 16) == 0) && (!x || ((uintptr_t)x % 16) == 0) && ((uintptr_t)gamma % 16) == 0 &&
                  ((uintptr_t)beta % 16) == 0, "gemm_ln: vector alignment");
  CUtensorMap ta, tb, tz, to, to16, tx;
  int rc;
  if ((rc = encode_2d(&ta, A, K, M, lda, BK, BM))) return rc;
  if ((rc = encode_2d(&tb, W, K, N, ldb, BK, PN / 2))) return rc;
  if ((rc = encode_2d_f32(&tz, z, N, M, N, 32, 32))) return rc;
  if ((rc = encode_2d_f32(&to, out, N, M, N, 32, 32))) return rc;
  if ((rc = encode_2d(&to16, out_bf16, N, M, N, 64, 32))) return rc;
  if ((rc = encode_2d_f32(&tx, x ? x : out, N, M, N, 32, 32))) return rc;
  LnArgs ep = {};
  ep.M = M; ep.N = N; ep.K = K; ep.kb_total = ceil_div(K, BK);
  ep.bias = bias; ep.x = x; ep.gamma = gamma; ep.beta = beta; ep.eps = eps; ep.mean = mean; ep.sigma = sigma;
  ep.use_drop = (p > 0.f && rng_state) ? 1 : 0;
  ep.drop.state = rng_state; ep.drop.salt = salt;
  ep.drop.thresh = (unsigned)(p * 65536.f + 0.5f); ep.drop.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  cudaStream_t s = (cudaStream_t)stream;
  return N == 256 ? launch_ln<1>(ta, tb, tz, to, to16, tx, ep, s) : launch_ln<2>(ta, tb, tz, to, to16, tx, ep, s);
}
