// Attention core on the 5th-gen tensor cores (bf16 arm).  One CTA (128 threads) per (sample, head):
//   forward   S = Q K^T          (tcgen05.mma 128 x N1 x 64, accumulator in TMEM columns [0,128))
//             softmax: thread i owns query row i == TMEM lane i, so row max / sum are thread-local (no shuffles);
//             the logits are re-read from TMEM in 32-column chunks by rolled loops (max pass, exp/sum pass): a
//             register-resident row needed 196 registers and 150-214 KB of straight-line code, which ran at the
//             speed of instruction-cache misses (ncu: 45 % of stalls no_instruction); chunked code is ~6x smaller,
//             fits 128 registers and lets the forward run four CTAs per SM (P aliases the dead Q/K tiles, O the
//             dead S columns), i.e. all 512 (sample, head) CTAs of a launch in one wave
//             P (bf16, dropout applied) -> shared memory in the canonical K-major 128B-swizzled UMMA layout
//             O = P V            (A = P K-major, B = V as an MN-major operand straight from its TMA tile)
//   backward  S = Q K^T, dP = dO V^T                       (V's tile re-read as a K-major operand)
//             dS = P (m dP - delta), delta_i = dO_i . O_i = sum_j P_ij m_ij dP_ij (fp32, from TMEM);  Pd, dS -> shared memory once, as [i][j] bf16
//             dV = Pd^T dO, dK = dS^T Q   (the same tiles read as MN-major A operands: no transposes)
//             dQ = dS K
// Q, K, V, dO tiles (128 rows x 64 head columns) arrive by TMA directly from the fused projection buffers
// (row pitch = buffer width); rows past Nq / Nk belong to the next sample or are zero-filled and are neutralised
// in the softmax (p = 0) — they never reach memory.  Semantics identical to attention.cu (reference
// modules.py:190-199, :232-240): masked_fill(-1e9) after the bias add, fully padded rows are uniform.
#include "tc_common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int TILE_BYTES = 128 * 64 * 2;        // one 128 x 64 bf16 operand tile
constexpr int PTILE_BYTES = 128 * 128 * 2;      // P / dS: two 64-wide k-chunks of 128 rows x 128 B
constexpr float LOG2E = 1.4426950408889634f;

struct AttnTcArgs {
  int B, heads, Nq, Nk;
  const unsigned char* kmask;
  const float* bias;
  __nv_bfloat16* o; long ldo;
  float scale;
  DropCfg drop;
  const __nv_bfloat16* dout; long lddo;
  __nv_bfloat16 *dq, *dk, *dv; long lddq, lddk, lddv;
  float* dbias;
};

// byte offset of element (row, col) in a [128 x 128] bf16 tile stored as two K-major SW128 chunks
__device__ __forceinline__ uint32_t p_offset(int row, int col8) {   // col8 = column / 8 (16-byte granule index, 0..15)
  const int kc = col8 >> 3, g = col8 & 7;
  return (uint32_t)(kc * TILE_BYTES + row * 128 + ((g ^ (row & 7)) << 4));
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Row-resident softmax state of one thread (= one query row = one TMEM lane): up to 4 chunks of 32 logits.
struct RowCtx {
  uint32_t mk0, mk1, mk2, mk3;   // padded-key bitmask, bit t of word c = key 32c+t (warp-uniform)
  const float* brow;       // bias row or null
  uint64_t rowbase;        // dropout element index of (row, key 0)
};

__device__ __forceinline__ RowCtx make_row_ctx(const AttnTcArgs& a, int b, int h, int i, bool row_ok, int lane) {
  RowCtx c;
  const unsigned char* mrow = a.kmask ? a.kmask + (size_t)b * a.Nk : nullptr;
  uint32_t mk[4];
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int j = w * 32 + lane;
    const bool m = mrow != nullptr && j < a.Nk && mrow[j] != 0;
    mk[w] = __ballot_sync(0xffffffffu, m);
  }
  c.mk0 = mk[0]; c.mk1 = mk[1]; c.mk2 = mk[2]; c.mk3 = mk[3];
  c.rowbase = (((uint64_t)b * a.heads + h) * a.Nq + i) * a.Nk;
  c.brow = (a.bias && row_ok) ? a.bias + c.rowbase : nullptr;
  return c;
}

__device__ __forceinline__ uint32_t chunk_mask(const RowCtx& c, int cc) {     // cc is a run-time loop index
  return cc < 2 ? (cc == 0 ? c.mk0 : c.mk1) : (cc == 2 ? c.mk2 : c.mk3);
}

// bias values of chunk cc (zeros without a bias / past Nk): issued BEFORE the TMEM loads of the chunk, so the global
// and the TMEM latencies overlap instead of adding up (ncu: 43 % of the stall samples were long-scoreboard waits)
template <bool VEC>
__device__ __forceinline__ void chunk_bias(const AttnTcArgs& a, const RowCtx& c, int cc, float (&bb)[32]) {
  const int j0 = cc * 32;
#pragma unroll
  for (int t = 0; t < 32; ++t) bb[t] = 0.f;
  if (c.brow) {
    if (VEC) {
#pragma unroll
      for (int g = 0; g < 8; ++g)
        if (j0 + 4 * g < a.Nk) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(c.brow + j0) + g);
          bb[4 * g] = v.x; bb[4 * g + 1] = v.y; bb[4 * g + 2] = v.z; bb[4 * g + 3] = v.w;
        }
    } else {
#pragma unroll
      for (int t = 0; t < 32; ++t)
        if (j0 + t < a.Nk) bb[t] = __ldg(c.brow + j0 + t);
    }
  }
}

// logits of chunk cc from the raw accumulator: scale, bias, mask; keys >= Nk -> -inf.
// VEC: bias / dbias rows are 16-byte aligned and dropout groups do not straddle rows (Nk % 4 == 0).
__device__ __forceinline__ void chunk_logits(const AttnTcArgs& a, const RowCtx& c, int cc, const uint32_t (&r)[32],
                                             const float (&bb)[32], float (&s)[32]) {
  const int j0 = cc * 32;
  const uint32_t mk = chunk_mask(c, cc);
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    float v = fmaf(__uint_as_float(r[t]), a.scale, bb[t]);
    if ((mk >> t) & 1u) v = -1e9f;
    s[t] = (j0 + t < a.Nk) ? v : -INFINITY;
  }
}

// Dropout decisions of the 32 keys [j0, j0 + 32) of this thread's row as one bit per key (1 = kept).  The stream is the
// library's: element e = rowbase + j of the flattened [B, heads, Nq, Nk] map takes 16 bits of hash(key, e >> 2).  A row
// that does not start on a multiple of 4 (Nk % 4 != 0: the 14 / 15 / 50-key shapes) shares its first and last hash with
// its neighbours, so a chunk needs 9 hashes instead of 8 — not one per element, which is what made the unaligned
// shapes 3-4x more expensive than the aligned ones.
__device__ __forceinline__ uint32_t chunk_keep_bits(const AttnTcArgs& a, const RowCtx& c, uint64_t key, int j0) {
  const uint64_t e0 = c.rowbase + (uint64_t)j0;
  const uint64_t g0 = e0 >> 2;
  const int o = (int)(e0 & 3);                 // position of column j0 inside its hash group
  uint32_t bits = 0;
#pragma unroll
  for (int g = 0; g < 9; ++g) {
    if (g == 8 && o == 0) break;
    const uint64_t r = mmnas_mix64(key ^ ((g0 + g) * 0x9E3779B97F4A7C15ull));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = 4 * g + u - o;             // column offset inside the chunk
      const uint32_t kept = (((unsigned)(r >> (16 * u)) & 0xFFFFu) < a.drop.thresh) ? 0u : 1u;
      if (t >= 0 && t < 32) bits |= kept << t;
    }
  }
  return bits;
}

__device__ __forceinline__ void tc_prologue(uint32_t bar_base, int nbars, uint32_t* tmem_slot, uint32_t cols, int warp) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < nbars; ++i) mbar_init(bar_base + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

template <bool VEC>
__global__ void __launch_bounds__(128, 4)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, AttnTcArgs a) {
  // Four CTAs per SM: 48 KB of shared memory and 128 TMEM columns each.
  //   smem: Q | K | V ; P (two 64-wide chunks) overwrites Q | K once S = Q K^T has retired
  //   TMEM: S [0,128), then O [0,64) once every thread has read its logits
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by POINTER arithmetic on the shared array: an integer round trip loses the address space
  // and every staging access becomes a generic ST.E / LD.E instead of STS / LDS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sQ = smem_u32(smem), sK = sQ + TILE_BYTES, sV = sK + TILE_BYTES, sP = sQ;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * TILE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const uint32_t bar = smem_u32(bars);
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int Nq = a.Nq, Nk = a.Nk;
  tc_prologue(bar, 3, tmem_slot, 128, warp);
  pdl_wait();                       // prologue above overlapped the previous kernel's tail
  const uint32_t tmem = *tmem_slot;
  const int n1 = max(16, (Nk + 15) & ~15);            // UMMA N of S = Q K^T
  if (tid == 0) {
    mbar_expect_tx(bar, 3 * TILE_BYTES);
    tma_load_2d(sQ, &tq, bar, h * 64, b * Nq);
    tma_load_2d(sK, &tk, bar, h * 64, b * Nk);
    tma_load_2d(sV, &tv, bar, h * 64, b * Nk);
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = make_idesc(false, false, 128, n1);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      umma_bf16(tmem, make_smem_desc(sQ + kk * 32, 16, 1024), make_smem_desc(sK + kk * 32, 16, 1024), idesc, kk > 0);
    umma_commit(bar + 8);
  }
  const int i = tid;
  const bool row_ok = i < Nq;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const int NC = (Nk + 31) >> 5;
  const bool use_drop = a.drop.state != nullptr && a.drop.thresh > 0;
  const uint64_t key = use_drop ? drop_key(a.drop) : 0;
  const RowCtx ctx = make_row_ctx(a, b, h, i, row_ok, tid & 31);
  mbar_wait(bar + 8, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_launch();

  float mx = -INFINITY;
#pragma unroll 1
  for (int cc = 0; cc < NC; ++cc) {                 // pass 1: row maximum
    uint32_t r[32];
    float s[32], bb[32];
    chunk_bias<VEC>(a, ctx, cc, bb);
    tmem_ld32(trow + cc * 32, r);
    chunk_logits(a, ctx, cc, r, bb, s);
#pragma unroll
    for (int t = 0; t < 32; ++t) mx = fmaxf(mx, s[t]);
  }
  float sum = 0.f;
  const float mxl = mx * LOG2E;
#pragma unroll 1
  for (int cc = 0; cc < NC; ++cc) {                 // pass 2: exp, row sum, dropout, P -> shared memory
    uint32_t r[32];
    float s[32], bb[32];
    chunk_bias<VEC>(a, ctx, cc, bb);
    tmem_ld32(trow + cc * 32, r);
    chunk_logits(a, ctx, cc, r, bb, s);
#pragma unroll
    for (int t = 0; t < 32; ++t) {
      s[t] = exp2f(fmaf(s[t], LOG2E, -mxl));        // -inf -> 0
      sum += s[t];
    }
    if (use_drop && row_ok) {
      const uint32_t kbits = chunk_keep_bits(a, ctx, key, cc * 32);
#pragma unroll
      for (int t = 0; t < 32; ++t) s[t] = ((kbits >> t) & 1u) ? s[t] * a.drop.scale : 0.f;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 pk = make_uint4(pack2(s[8 * g], s[8 * g + 1]), pack2(s[8 * g + 2], s[8 * g + 3]),
                            pack2(s[8 * g + 4], s[8 * g + 5]), pack2(s[8 * g + 6], s[8 * g + 7]));
      *reinterpret_cast<uint4*>(smem + p_offset(i, cc * 4 + g)) = pk;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy P writes -> visible to the MMA
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();                                               // every thread has read its S columns: O may overwrite them
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = make_idesc(false, true, 128, 64);
    const int ksteps = (Nk + 15) >> 4;
    for (int t = 0; t < ksteps; ++t)
      umma_bf16(tmem, make_smem_desc(sP + (t >> 2) * TILE_BYTES + (t & 3) * 32, 16, 1024),
                make_smem_desc(sV + t * 2048, 8192, 1024), idesc, t > 0);
    umma_commit(bar + 16);
  }
  mbar_wait(bar + 16, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll 1
  for (int cc = 0; cc < 2; ++cc) {
    uint32_t r[32];
    tmem_ld32(trow + cc * 32, r);
    if (row_ok) {
      __nv_bfloat16* orow = a.o + ((size_t)b * Nq + i) * a.ldo + h * 64 + cc * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 pk = make_uint4(pack2(__uint_as_float(r[8 * g]) * inv, __uint_as_float(r[8 * g + 1]) * inv),
                              pack2(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv),
                              pack2(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv),
                              pack2(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv));
        *reinterpret_cast<uint4*>(orow + 8 * g) = pk;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

template <bool VEC>
__global__ void __launch_bounds__(128, 2)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo, AttnTcArgs a) {
  // Two CTAs per SM: 112 KB of shared memory and 256 TMEM columns each.
  //   smem: Q | K | dO | P chunk 0 | V (= P chunk 1 once dP = dO V^T has retired) | dS chunk 0 | dS chunk 1
  //   TMEM: phase 1  S [0,128)  dP [128,256);  phase 2 (S, dP consumed)  dQ [0,64)  dK [64,128)  dV [128,192)
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem), sK = sQ + TILE_BYTES, sdO = sK + TILE_BYTES, sP = sdO + TILE_BYTES;
  const uint32_t sV = sP + TILE_BYTES, sdS = sV + TILE_BYTES;
  uint8_t* gP = smem + 3 * TILE_BYTES;
  uint8_t* gdS = gP + PTILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * TILE_BYTES + 2 * PTILE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const uint32_t bar = smem_u32(bars);
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int Nq = a.Nq, Nk = a.Nk;
  tc_prologue(bar, 3, tmem_slot, 256, warp);
  pdl_wait();
  const uint32_t tmem = *tmem_slot;
  const int n1 = max(16, (Nk + 15) & ~15);
  if (tid == 0) {
    mbar_expect_tx(bar, 4 * TILE_BYTES);
    tma_load_2d(sQ, &tq, bar, h * 64, b * Nq);
    tma_load_2d(sK, &tk, bar, h * 64, b * Nk);
    tma_load_2d(sV, &tv, bar, h * 64, b * Nk);
    tma_load_2d(sdO, &tdo, bar, h * 64, b * Nq);
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = make_idesc(false, false, 128, n1);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      umma_bf16(tmem, make_smem_desc(sQ + kk * 32, 16, 1024), make_smem_desc(sK + kk * 32, 16, 1024), idesc, kk > 0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)      // dP = dO V^T : V's tile as a K-major operand
      umma_bf16(tmem + 128, make_smem_desc(sdO + kk * 32, 16, 1024), make_smem_desc(sV + kk * 32, 16, 1024), idesc, kk > 0);
    umma_commit(bar + 8);
  }
  const int i = tid;
  const bool row_ok = i < Nq;
  mbar_wait(bar + 8, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const int NC = (Nk + 31) >> 5;
  const bool use_drop = a.drop.state != nullptr && a.drop.thresh > 0;
  const uint64_t key = use_drop ? drop_key(a.drop) : 0;
  const RowCtx ctx = make_row_ctx(a, b, h, i, row_ok, tid & 31);
  // pass 1: row maximum, row sum and delta_i = dO_i . O_i = sum_j P_ij m_ij dP_ij together (running rescale).
  // delta is formed in fp32 from the TMEM-resident S and dP, i.e. from the SAME probabilities the dS pass uses, so
  // sum_j dS_ij cancels to fp32 rounding.  (Taking it from the bf16-rounded O and dO in global memory left a
  // residue of ~2^-9 |delta| in every row sum, which is what d linear_r.bias = sum dS / r accumulates.)
  float mx = -INFINITY, sum = 0.f, num = 0.f;
  uint32_t kb0 = 0xFFFFFFFFu, kb1 = 0xFFFFFFFFu, kb2 = 0xFFFFFFFFu, kb3 = 0xFFFFFFFFu;    // dropout keep bits, hashed once
#pragma unroll 1
  for (int cc = 0; cc < NC; ++cc) {
    uint32_t r[32], rp[32];
    float s[32], bb[32];
    chunk_bias<VEC>(a, ctx, cc, bb);
    tmem_ld32_nowait(trow + cc * 32, r);
    tmem_ld32_nowait(trow + 128 + cc * 32, rp);
    tmem_ld_wait();
    chunk_logits(a, ctx, cc, r, bb, s);
    float cm = s[0];
#pragma unroll
    for (int t = 1; t < 32; ++t) cm = fmaxf(cm, s[t]);
    const float nm = fmaxf(mx, cm);                 // finite from chunk 0 on: key 0 is never -inf
    const float nml = nm * LOG2E;
    float part = 0.f, pnum = 0.f;
    const uint32_t kbits = use_drop ? chunk_keep_bits(a, ctx, key, cc * 32) : 0xFFFFFFFFu;
    const float dscale1 = use_drop ? a.drop.scale : 1.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float m[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) m[u] = ((kbits >> (4 * g + u)) & 1u) ? dscale1 : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = 4 * g + u;
        const float e = exp2f(fmaf(s[t], LOG2E, -nml));
        part += e;
        // dP columns past Nk were never written by the MMA (TMEM garbage, possibly NaN): select, do not multiply
        const float dp = (cc * 32 + t < Nk) ? __uint_as_float(rp[t]) : 0.f;
        pnum = fmaf(e * m[u], dp, pnum);
      }
    }
    // rescale by exactly the ratio of the two term scalings (1 when the maximum did not move, even at -1e9 where
    // the rounded products carry errors of +-64; 0 on the first chunk, mx = -inf)
    const float f = exp2f(__fmul_rn(mx, LOG2E) - nml);             // __fmul_rn: no contraction into an fma
    sum = fmaf(sum, f, part);
    num = fmaf(num, f, pnum);
    mx = nm;
    if (cc < 2) { if (cc == 0) kb0 = kbits; else kb1 = kbits; } else { if (cc == 2) kb2 = kbits; else kb3 = kbits; }
  }
  const float mxl = mx * LOG2E;
  const float inv = (row_ok && sum > 0.f) ? 1.f / sum : 0.f;     // rows >= Nq contribute zeros to dK / dV
  const float delta = num * inv;
  float* dbrow = (a.dbias && row_ok) ? a.dbias + ctx.rowbase : nullptr;
#pragma unroll 1
  for (int cc = 0; cc < NC; ++cc) {                 // pass 2: P (dropped), dS -> shared memory; d bias -> global
    uint32_t r[32], rp[32];
    float s[32], ds[32], bb[32];
    chunk_bias<VEC>(a, ctx, cc, bb);
    tmem_ld32_nowait(trow + cc * 32, r);
    tmem_ld32_nowait(trow + 128 + cc * 32, rp);
    tmem_ld_wait();
    chunk_logits(a, ctx, cc, r, bb, s);
    const uint32_t mk = chunk_mask(ctx, cc);
    const uint32_t kbits = cc < 2 ? (cc == 0 ? kb0 : kb1) : (cc == 2 ? kb2 : kb3);      // the masks pass 1 hashed
    const float dscale = use_drop ? a.drop.scale : 1.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float m[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) m[u] = ((kbits >> (4 * g + u)) & 1u) ? dscale : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = 4 * g + u;
        const float p = exp2f(fmaf(s[t], LOG2E, -mxl)) * inv;    // 0 for keys >= Nk (exp2(-inf)) and rows >= Nq
        float d = p * (m[u] * __uint_as_float(rp[t]) - delta);
        if (((mk >> t) & 1u) || p == 0.f) d = 0.f;               // masked_fill cuts the graph; p == 0 guards garbage dP
        s[t] = p * m[u];
        ds[t] = d;
      }
      if (dbrow && cc * 32 + 4 * g < Nk) {
        if (VEC) {
          *reinterpret_cast<float4*>(dbrow + cc * 32 + 4 * g) = make_float4(ds[4 * g], ds[4 * g + 1], ds[4 * g + 2], ds[4 * g + 3]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (cc * 32 + 4 * g + u < Nk) dbrow[cc * 32 + 4 * g + u] = ds[4 * g + u];
        }
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t off = p_offset(i, cc * 4 + g);
      *reinterpret_cast<uint4*>(gP + off) = make_uint4(pack2(s[8 * g], s[8 * g + 1]), pack2(s[8 * g + 2], s[8 * g + 3]),
                                                       pack2(s[8 * g + 4], s[8 * g + 5]), pack2(s[8 * g + 6], s[8 * g + 7]));
      *reinterpret_cast<uint4*>(gdS + off) = make_uint4(pack2(ds[8 * g], ds[8 * g + 1]), pack2(ds[8 * g + 2], ds[8 * g + 3]),
                                                        pack2(ds[8 * g + 4], ds[8 * g + 5]), pack2(ds[8 * g + 6], ds[8 * g + 7]));
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int ksteps_j = (Nk + 15) >> 4, ksteps_i = (Nq + 15) >> 4;
    // dQ = dS K : A = dS K-major (k = j), B = K tile as MN-major (k-rows j, n = head column)
    const uint32_t id_q = make_idesc(false, true, 128, 64);
    for (int t = 0; t < ksteps_j; ++t)
      umma_bf16(tmem + 0, make_smem_desc(sdS + (t >> 2) * TILE_BYTES + (t & 3) * 32, 16, 1024),
                make_smem_desc(sK + t * 2048, 8192, 1024), id_q, t > 0);
    // dK = dS^T Q, dV = Pd^T dO : A = the same [i][j] tiles read MN-major (m = j: two 64-wide chunks, k-rows = i)
    const uint32_t id_t = make_idesc(true, true, 128, 64);
    for (int t = 0; t < ksteps_i; ++t)
      umma_bf16(tmem + 64, make_smem_desc(sdS + t * 2048, TILE_BYTES, 1024), make_smem_desc(sQ + t * 2048, 8192, 1024), id_t, t > 0);
    for (int t = 0; t < ksteps_i; ++t)
      umma_bf16(tmem + 128, make_smem_desc(sP + t * 2048, TILE_BYTES, 1024), make_smem_desc(sdO + t * 2048, 8192, 1024), id_t, t > 0);
    umma_commit(bar + 16);
  }
  mbar_wait(bar + 16, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_launch();
  // epilogue: thread t owns dQ row i = t and dK / dV row j = t
#pragma unroll 1
  for (int which = 0; which < 3; ++which) {
    const bool ok = which == 0 ? row_ok : (tid < Nk);
    const float mul = which == 2 ? 1.f : a.scale;
    __nv_bfloat16* base = which == 0 ? a.dq + ((size_t)b * Nq + tid) * a.lddq
                        : which == 1 ? a.dk + ((size_t)b * Nk + tid) * a.lddk
                                     : a.dv + ((size_t)b * Nk + tid) * a.lddv;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t r[32];
      tmem_ld32(trow + which * 64 + cc * 32, r);
      if (ok) {
        __nv_bfloat16* row = base + h * 64 + cc * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk = make_uint4(pack2(__uint_as_float(r[8 * g]) * mul, __uint_as_float(r[8 * g + 1]) * mul),
                                pack2(__uint_as_float(r[8 * g + 2]) * mul, __uint_as_float(r[8 * g + 3]) * mul),
                                pack2(__uint_as_float(r[8 * g + 4]) * mul, __uint_as_float(r[8 * g + 5]) * mul),
                                pack2(__uint_as_float(r[8 * g + 6]) * mul, __uint_as_float(r[8 * g + 7]) * mul));
          *reinterpret_cast<uint4*>(row + 8 * g) = pk;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

constexpr int FWD_SMEM = 3 * TILE_BYTES + 1024 + 128;
constexpr int BWD_SMEM = 3 * TILE_BYTES + 2 * PTILE_BYTES + 128;

DropCfg mk_drop(const unsigned long long* st, unsigned long long salt, float p) {
  DropCfg d;
  d.state = p > 0.f ? st : nullptr; d.salt = salt;
  d.thresh = (unsigned)(p * 65536.f + 0.5f); d.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  return d;
}

bool aligned16(const void* p, long ld) { return ((uintptr_t)p % 16) == 0 && (ld % 8) == 0; }

}  // namespace

// Returns MMNAS_ERR_UNSUPPORTED when the operands do not meet the TMA constraints (caller then uses the FFMA kernel).
int mmnas_attn_fwd_tc(int B, int heads, int Nq, int Nk, const void* q, long ldq, const void* k, long ldk, const void* v,
                      long ldv, const unsigned char* kmask, const float* bias, void* o, long ldo, float scale,
                      const unsigned long long* rng_state, unsigned long long salt, float p, cudaStream_t s) {
  if (Nq > 128 || Nk > 128 || !aligned16(q, ldq) || !aligned16(k, ldk) || !aligned16(v, ldv) || !aligned16(o, ldo))
    return MMNAS_ERR_UNSUPPORTED;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = encode_2d(&tq, q, heads * 64, (long)B * Nq, ldq, 64, 128))) return rc;
  if ((rc = encode_2d(&tk, k, heads * 64, (long)B * Nk, ldk, 64, 128))) return rc;
  if ((rc = encode_2d(&tv, v, heads * 64, (long)B * Nk, ldv, 64, 128))) return rc;
  AttnTcArgs a = {};
  a.B = B; a.heads = heads; a.Nq = Nq; a.Nk = Nk; a.kmask = kmask; a.bias = bias;
  a.o = (__nv_bfloat16*)o; a.ldo = ldo; a.scale = scale; a.drop = mk_drop(rng_state, salt, p);
  static bool attr_done = false;
  if (!attr_done) {
    MMNAS_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    MMNAS_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    attr_done = true;
  }
  if ((Nk & 3) == 0) MMNAS_CUDA(mmnas_launch(attn_fwd_tc_kernel<true>, dim3(heads, B), dim3(128), FWD_SMEM, s, tq, tk, tv, a));
  else MMNAS_CUDA(mmnas_launch(attn_fwd_tc_kernel<false>, dim3(heads, B), dim3(128), FWD_SMEM, s, tq, tk, tv, a));
  return MMNAS_OK;
}

int mmnas_attn_bwd_tc(int B, int heads, int Nq, int Nk, const void* q, long ldq, const void* k, long ldk, const void* v,
                      long ldv, const unsigned char* kmask, const float* bias, const void* o, long ldo, const void* dout,
                      long lddo, void* dq, long lddq, void* dk, long lddk, void* dv, long lddv, float* dbias, float scale,
                      const unsigned long long* rng_state, unsigned long long salt, float p, cudaStream_t s) {
  if (Nq > 128 || Nk > 128 || !aligned16(q, ldq) || !aligned16(k, ldk) || !aligned16(v, ldv) || !aligned16(o, ldo) ||
      !aligned16(dout, lddo) || !aligned16(dq, lddq) || !aligned16(dk, lddk) || !aligned16(dv, lddv))
    return MMNAS_ERR_UNSUPPORTED;
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = encode_2d(&tq, q, heads * 64, (long)B * Nq, ldq, 64, 128))) return rc;
  if ((rc = encode_2d(&tk, k, heads * 64, (long)B * Nk, ldk, 64, 128))) return rc;
  if ((rc = encode_2d(&tv, v, heads * 64, (long)B * Nk, ldv, 64, 128))) return rc;
  if ((rc = encode_2d(&tdo, dout, heads * 64, (long)B * Nq, lddo, 64, 128))) return rc;
  AttnTcArgs a = {};
  a.B = B; a.heads = heads; a.Nq = Nq; a.Nk = Nk; a.kmask = kmask; a.bias = bias;
  a.o = (__nv_bfloat16*)const_cast<void*>(o); a.ldo = ldo; a.scale = scale; a.drop = mk_drop(rng_state, salt, p);
  a.dout = (const __nv_bfloat16*)dout; a.lddo = lddo;
  a.dq = (__nv_bfloat16*)dq; a.dk = (__nv_bfloat16*)dk; a.dv = (__nv_bfloat16*)dv;
  a.lddq = lddq; a.lddk = lddk; a.lddv = lddv; a.dbias = dbias;
  static bool attr_done = false;
  if (!attr_done) {
    MMNAS_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    MMNAS_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    attr_done = true;
  }
  if ((Nk & 3) == 0) MMNAS_CUDA(mmnas_launch(attn_bwd_tc_kernel<true>, dim3(heads, B), dim3(128), BWD_SMEM, s, tq, tk, tv, tdo, a));
  else MMNAS_CUDA(mmnas_launch(attn_bwd_tc_kernel<false>, dim3(heads, B), dim3(128), BWD_SMEM, s, tq, tk, tv, tdo, a));
  return MMNAS_OK;
}
