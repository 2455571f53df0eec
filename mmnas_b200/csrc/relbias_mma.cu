// RSA geometry bias on the tensor cores (bf16 arm, geometry mode, 2 / 4 / 6 / 8 heads: H = 512 train, H = 256 search) — the warp-level companion of relbias.cu.
//
//   e_c   = relu(W_y[c,:] . g + b_y[c])                 64 channels per (i,j) pair      (full_vqa.py:82,103)
//   r_h   = W_r[h,:] . e + b_r[h]                        8 heads                         (modules.py:231)
//   bias  = log(max(relu(r), 1e-6))                                                      (modules.py:235)
//   bwd:  d pre_r = dbias / r (where r > 1e-6),  d e = W_r^T d pre_r (through the relu of e),
//         dW_r += d pre_r (x) e,  dW_y += d e (x) g,  db_y += d e,  db_r += d pre_r      (summed over all pairs)
//
// Of the ~2 100 multiply-adds per pair of the backward only the 256 of the first layer are left on the FP32 pipe
// (relbias.cu is bound by it: three-register FFMA / FFMA2 issue at half rate on sm_100).  Everything with the 64-wide
// channel axis is a small matrix product and runs as mma.sync.m16n8k16 (bf16 in, fp32 accumulate) on REGISTER
// fragments — no shared-memory tiles, no TMEM, no barriers inside the loop:
//   * a warp owns 32 pairs per iteration = two 16-row M tiles.  Thread (g = lane/4, q = lane%4) computes e for pairs
//     {g, g+8, g+16, g+24} and channels {8n + 2q, 8n + 2q + 1 : n = 0..7}, which IS the A-fragment layout of
//     r = e W_r^T (K = channels); W_r sits in registers as B fragments.  e and W_r enter as hi + lo bf16 pairs
//     (hi.hi + lo.hi + hi.lo), so r — whose logarithm becomes an attention logit — keeps ~16 mantissa bits;
//   * the accumulator fragment of r (rows g, g+8; columns 2q, 2q+1) is exactly the A fragment (K = heads, upper half
//     zero) of d e = d pre_r W_r, and its output fragment lands on the same (pair, channel) set the thread computed e
//     for, so the relu mask is thread-local;
//   * the reductions over pairs (dW_r^T = e^T d pre_r, dW_y^T = d e^T [g | 1]) contract over the fragment ROW index:
//     movmatrix.m8n8.trans flips the 8x8 register tiles in place, and the products accumulate in 32 registers per
//     thread across the whole persistent loop; one shared-memory reduction and one global atomic per output per CTA
//     at the end.
// The fp32 arm, the dense-rel_embed compatibility mode and other head counts stay on relbias.cu.
#include <cuda_bf16.h>
#include "common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int R = 64;
constexpr int HEADS = 8;
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;

struct MmaArgs {
  unsigned pairs, nn;
  int heads;                 // even, <= 8: the head axis is the n = 8 (or k = 8 of 16) axis of the fragments, zero padded
  const float *g4, *Wy, *by, *Wr, *br;
  float* bias;
  const float* dbias;
  float *dWy, *dby, *dWr, *dbr;
};

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// (x0, x1) -> bf16 pair `hi` and the bf16 pair of the rounding residuals `lo` (one packed convert each way)
__device__ __forceinline__ void split_bf16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(x0, x1);
  lo = pack_bf16(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xFFFF0000u));
}
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// transpose an 8x8 tile of 16-bit elements held one packed pair per lane (row lane/4, columns 2(lane%4), +1)
__device__ __forceinline__ uint32_t tile_t(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

// MT = 16-pair M tiles per warp iteration: 2 in the forward; 1 in the backward, whose per-tile state (e fragments,
// d e, geometry) would otherwise push it past 128 registers and down to one CTA per SM.
template <bool BWD, int MT>
__global__ void __launch_bounds__(THREADS, 2) relbias_mma_kernel(MmaArgs a) {
  __shared__ float4 sWy[R];            // W_y[c][0..3]
  __shared__ float sby[R];
  __shared__ float sred[R * HEADS + R * 4 + R + HEADS];     // dW_r | dW_y | db_y | db_r of this CTA
  pdl_wait(); pdl_launch();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  for (int i = tid; i < R; i += THREADS) {
    sWy[i] = __ldg(reinterpret_cast<const float4*>(a.Wy) + i);
    sby[i] = __ldg(a.by + i);
  }
  if (BWD)
    for (int i = tid; i < R * HEADS + R * 4 + R + HEADS; i += THREADS) sred[i] = 0.f;
  // B fragments of r = e W_r^T: k-step k covers channels 16k..16k+15; b0 rows 2q, 2q+1, b1 rows 2q+8, 2q+9; column = head g
  uint32_t wr_hi[4][2], wr_lo[4][2];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float* w = a.Wr + g * R + 16 * k + 2 * q;
    const bool hv = g < a.heads;                   // heads past `heads`: zero columns
    split_bf16(hv ? __ldg(w) : 0.f, hv ? __ldg(w + 1) : 0.f, wr_hi[k][0], wr_lo[k][0]);
    split_bf16(hv ? __ldg(w + 8) : 0.f, hv ? __ldg(w + 9) : 0.f, wr_hi[k][1], wr_lo[k][1]);
  }
  // B fragments of d e = d pre_r W_r: K = heads (8, padded to 16), n-tile n covers channels 8n..8n+7; column = channel 8n+g
  uint32_t wde[8];
  if (BWD) {
#pragma unroll
    for (int n = 0; n < 8; ++n)
      wde[n] = 2 * q < a.heads ? pack_bf16(__ldg(a.Wr + (2 * q) * R + 8 * n + g), __ldg(a.Wr + (2 * q + 1) * R + 8 * n + g)) : 0u;
  }
  const bool myheads = 2 * q < a.heads;            // this thread's head pair (2q, 2q+1) exists
  const float br0 = myheads ? __ldg(a.br + 2 * q) : 0.f, br1 = myheads ? __ldg(a.br + 2 * q + 1) : 0.f;
  float accWr[4][4], accWy[4][4], accbr0 = 0.f, accbr1 = 0.f;
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int i = 0; i < 4; ++i) { accWr[m][i] = 0.f; accWy[m][i] = 0.f; }
  __syncthreads();

  constexpr unsigned PPI = 16u * MT;           // pairs per warp iteration
  const unsigned iters = (a.pairs + PPI - 1u) / PPI;
  const unsigned stride = gridDim.x * WARPS;
  // geometry of the NEXT iteration's four pairs is fetched while the current one is processed
  float4 gnext[MT][2];
  auto fetch_g = [&](unsigned wi_) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const unsigned p = wi_ * PPI + 16 * mt + 8 * rh + g;
        gnext[mt][rh] = (wi_ < iters && p < a.pairs) ? __ldg(reinterpret_cast<const float4*>(a.g4) + p) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
  };
  fetch_g(blockIdx.x * WARPS + warp);
  for (unsigned wi = blockIdx.x * WARPS + warp; wi < iters; wi += stride) {
    const unsigned base = wi * PPI;
    // this thread's pairs: M tile mt, row half rh -> base + 16 mt + 8 rh + g.  The pairs of an iteration
    // straddle few image boundaries: one integer division per iteration, then a short walk.
    const unsigned bb = base / a.nn, ij_base = base - bb * a.nn;
    bool live[MT][2];
    size_t off[MT][2];                    // element offset of (pair, head 2q) in bias / dbias
    float4 gv[MT][2];
    float db[MT][2][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const unsigned d = 16 * mt + 8 * rh + g;
        live[mt][rh] = base + d < a.pairs && myheads;
        unsigned b = bb, ij = ij_base + d;
        while (ij >= a.nn) { ij -= a.nn; ++b; }
        off[mt][rh] = ((size_t)b * a.heads + 2 * q) * a.nn + ij;
        gv[mt][rh] = gnext[mt][rh];
        if (BWD) {        // issued early: latency hides behind the first layer
          db[mt][rh][0] = live[mt][rh] ? __ldg(a.dbias + off[mt][rh]) : 0.f;
          db[mt][rh][1] = live[mt][rh] ? __ldg(a.dbias + off[mt][rh] + a.nn) : 0.f;
        }
      }
    fetch_g(wi + stride);
    // ---- first layer on the FP32 pipe, straight into A-fragment registers (hi / lo bf16 pairs), and
    // ---- r = e W_r^T + b_r per M tile as soon as a 16-channel k-step is complete: accumulator rows g (rh 0) and
    // ---- g+8 (rh 1), columns = heads 2q, 2q+1.  The lo halves live for one k-step only.
    uint32_t e_hi[MT][2][8];
    float rr[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) { rr[mt][0] = br0; rr[mt][1] = br1; rr[mt][2] = br0; rr[mt][3] = br1; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t e_lo[MT][2][2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = 2 * k + j;
        const int c = 8 * n + 2 * q;
        const float4 w0 = sWy[c], w1 = sWy[c + 1];
        const float b0 = sby[c], b1 = sby[c + 1];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
            const float4 v = gv[mt][rh];
            const float x0 = fmaxf(fmaf(w0.x, v.x, fmaf(w0.y, v.y, fmaf(w0.z, v.z, fmaf(w0.w, v.w, b0)))), 0.f);
            const float x1 = fmaxf(fmaf(w1.x, v.x, fmaf(w1.y, v.y, fmaf(w1.z, v.z, fmaf(w1.w, v.w, b1)))), 0.f);
            split_bf16(x0, x1, e_hi[mt][rh][n], e_lo[mt][rh][j]);
          }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        mma16816(rr[mt], e_hi[mt][0][2 * k], e_hi[mt][1][2 * k], e_hi[mt][0][2 * k + 1], e_hi[mt][1][2 * k + 1], wr_hi[k][0], wr_hi[k][1]);
        mma16816(rr[mt], e_lo[mt][0][0], e_lo[mt][1][0], e_lo[mt][0][1], e_lo[mt][1][1], wr_hi[k][0], wr_hi[k][1]);
        mma16816(rr[mt], e_hi[mt][0][2 * k], e_hi[mt][1][2 * k], e_hi[mt][0][2 * k + 1], e_hi[mt][1][2 * k + 1], wr_lo[k][0], wr_lo[k][1]);
      }
    }
    if constexpr (!BWD) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int rh = 0; rh < 2; ++rh)
          if (live[mt][rh]) {
            float* o = a.bias + off[mt][rh];
            o[0] = logf(fmaxf(fmaxf(rr[mt][2 * rh], 0.f), 1e-6f));
            o[a.nn] = logf(fmaxf(fmaxf(rr[mt][2 * rh + 1], 0.f), 1e-6f));
          }
    } else {
    // ---- backward
    uint32_t dp[MT][2];                   // d pre_r as A-fragment rows (K = heads 2q, 2q+1)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const float r0 = rr[mt][2 * rh], r1 = rr[mt][2 * rh + 1];
        const float d0 = (live[mt][rh] && r0 > 1e-6f) ? __fdividef(db[mt][rh][0], r0) : 0.f;    // relu and clamp both pass
        const float d1 = (live[mt][rh] && r1 > 1e-6f) ? __fdividef(db[mt][rh][1], r1) : 0.f;
        accbr0 += d0; accbr1 += d1;
        dp[mt][rh] = pack_bf16(d0, d1);
      }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      // d e = d pre_r W_r, masked by e > 0, as bf16 A-fragment tiles [pair rows][channels]
      uint32_t de[2][8];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(d, dp[mt][0], dp[mt][1], 0u, 0u, wde[n], 0u);
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          const uint32_t eh = e_hi[mt][rh][n];
          de[rh][n] = pack_bf16((eh & 0xFFFFu) ? d[2 * rh] : 0.f, (eh >> 16) ? d[2 * rh + 1] : 0.f);
        }
      }
      // reductions over the 16 pairs of this M tile: transposed tiles as A (rows = channels), pairs as K
      const uint32_t dpt0 = tile_t(dp[mt][0]), dpt1 = tile_t(dp[mt][1]);     // B: rows = pairs, column = head g
      // B of dW_y: rows = pairs 2q, 2q+1 (b0) / 2q+8, 2q+9 (b1), column n = g: geometry component g, 1 for n == 4
      float gb[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const unsigned p0 = base + 16 * mt + 2 * q;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned p = p0 + (u & 1) + 8 * (u >> 1);
          if (g < 4) gb[u] = p < a.pairs ? __ldg(a.g4 + (size_t)p * 4 + g) : 0.f;
          else if (g == 4) gb[u] = 1.f;
        }
      }
      const uint32_t gq0 = pack_bf16(gb[0], gb[1]), gq1 = pack_bf16(gb[2], gb[3]);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        mma16816(accWr[m], tile_t(e_hi[mt][0][2 * m]), tile_t(e_hi[mt][0][2 * m + 1]), tile_t(e_hi[mt][1][2 * m]),
                 tile_t(e_hi[mt][1][2 * m + 1]), dpt0, dpt1);
        mma16816(accWy[m], tile_t(de[0][2 * m]), tile_t(de[0][2 * m + 1]), tile_t(de[1][2 * m]), tile_t(de[1][2 * m + 1]), gq0, gq1);
      }
    }
    }   // backward
  }
  if (!BWD) return;
  // ---- CTA reduction, then one global atomic per output.  Accumulator (m, i): channel 16m + g + 8 (i >> 1), column 2q + (i & 1)
  float* sWr = sred;
  float* sWyr = sred + R * HEADS;
  float* sbyr = sWyr + R * 4;
  float* sbr = sbyr + R;
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = 16 * m + g + 8 * (i >> 1), col = 2 * q + (i & 1);
      if (col < a.heads) atomicAdd(&sWr[col * R + c], accWr[m][i]);
      if (col < 4) atomicAdd(&sWyr[c * 4 + col], accWy[m][i]);
      else if (col == 4) atomicAdd(&sbyr[c], accWy[m][i]);
    }
  // d b_r: lanes with the same q hold partial sums of heads 2q, 2q+1
#pragma unroll
  for (int off = 4; off < 32; off <<= 1) {
    accbr0 += __shfl_xor_sync(0xffffffffu, accbr0, off);
    accbr1 += __shfl_xor_sync(0xffffffffu, accbr1, off);
  }
  if (g == 0 && myheads) { atomicAdd(&sbr[2 * q], accbr0); atomicAdd(&sbr[2 * q + 1], accbr1); }
  __syncthreads();
  for (int i = tid; i < R * a.heads; i += THREADS) atomicAdd(&a.dWr[i], sWr[i]);
  for (int i = tid; i < R * 4; i += THREADS) atomicAdd(&a.dWy[i], sWyr[i]);
  for (int i = tid; i < R; i += THREADS) atomicAdd(&a.dby[i], sbyr[i]);
  if (tid < a.heads) atomicAdd(&a.dbr[tid], sbr[tid]);
}

int grid_for(unsigned pairs, unsigned pairs_per_iter, unsigned ctas_per_sm) {
  const unsigned iters = (pairs + pairs_per_iter - 1u) / pairs_per_iter;
  const unsigned ctas = (iters + WARPS - 1) / WARPS;
  return (int)(ctas < 148u * ctas_per_sm ? ctas : 148u * ctas_per_sm);      // persistent: one resident wave
}

}  // namespace

// Returns MMNAS_ERR_UNSUPPORTED when the configuration is outside this kernel (caller then uses relbias.cu).
int mmnas_relbias_fwd_mma(int B, int N, int heads, const float* g4, const float* Wy, const float* by, const float* Wr,
                          const float* br, float* bias, cudaStream_t s) {
  if (heads > HEADS || heads < 2 || (heads & 1) || !g4 || ((uintptr_t)g4 % 16) != 0 || ((uintptr_t)Wy % 16) != 0) return MMNAS_ERR_UNSUPPORTED;
  MmaArgs a = {};
  a.nn = (unsigned)N * N; a.pairs = (unsigned)B * a.nn;
  a.heads = heads;
  a.g4 = g4; a.Wy = Wy; a.by = by; a.Wr = Wr; a.br = br; a.bias = bias;
  MMNAS_CUDA(mmnas_launch(relbias_mma_kernel<false, 2>, dim3(grid_for(a.pairs, 32, 2)), dim3(THREADS), 0, s, a));
  return MMNAS_OK;
}

int mmnas_relbias_bwd_mma(int B, int N, int heads, const float* g4, const float* Wy, const float* by, const float* Wr,
                          const float* br, const float* dbias, float* dWy, float* dby, float* dWr, float* dbr,
                          cudaStream_t s) {
  if (heads > HEADS || heads < 2 || (heads & 1) || !g4 || ((uintptr_t)g4 % 16) != 0 || ((uintptr_t)Wy % 16) != 0) return MMNAS_ERR_UNSUPPORTED;
  MmaArgs a = {};
  a.nn = (unsigned)N * N; a.pairs = (unsigned)B * a.nn;
  a.heads = heads;
  a.g4 = g4; a.Wy = Wy; a.by = by; a.Wr = Wr; a.br = br; a.dbias = dbias;
  a.dWy = dWy; a.dby = dby; a.dWr = dWr; a.dbr = dbr;
  MMNAS_CUDA(mmnas_launch(relbias_mma_kernel<true, 1>, dim3(grid_for(a.pairs, 16, 2)), dim3(THREADS), 0, s, a));
  return MMNAS_OK;
}
