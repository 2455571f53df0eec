// Question / caption encoder LSTM (reference: full_vqa.py:68-74,94-95 — nn.LSTM(WORD_EMBED_SIZE -> HSIZE, one layer,
// batch_first, zero initial state; only the output sequence is used).  bf16 arm only: the fp32 arm keeps torch's LSTM.
//
// The input projection X W_ih^T + b_ih + b_hh of all time steps is one GEMM (mmnas_gemm_bf16) done by the caller.
// What is left is the recurrence: T dependent steps of  gates_t = xw_t + h_{t-1} W_hh^T  ([B, 4H], K = H) and the cell
// update — 134 MFLOP per step at B = 64: far too little for a launch per step (cuDNN: two launches per step, ~8.5 us
// per step forward on a B200).  Batch rows are independent, the hidden units of one row are not, so:
//   * a CLUSTER of 16 CTAs owns a group of 16 (B <= 128) or 32 batch rows for the whole sequence; CTA r of the cluster
//     owns hidden units [r H/16, (r+1) H/16), i.e. H/4 of the 4H gate columns;
//   * its slice of W_hh (128 x 512 bf16 at H = 512) lives in REGISTERS as mma.sync B fragments for all T steps (128
//     registers per thread, 8 warps), so a step moves no weights at all;
//   * h_{t-1} of the row group (bf16, [rows, H]) sits in every CTA's shared memory; per step a CTA computes its
//     [rows x H/4] gate block with mma.sync m16n8k16 (fp32 accumulation).  The k index inside a 32-wide chunk is
//     permuted identically for A and B so that both fragments are single 16-byte loads;
//   * the four gates of one unit land in the four lanes of a quad (columns ordered [i i f f g g o o] per unit pair):
//     activations, quad shuffles, cell update in registers — the cell state never leaves its thread;
//   * the CTA's new h slice goes to the 16 CTAs' shared memory with st.async (remote stores that complete transaction
//     bytes on the RECEIVER's mbarrier, double-buffered): a CTA starts step t+1 as soon as the 16 slices of h_t have
//     landed — no cluster-wide barrier, no release fence on the sender, nothing in global memory to spin on.
// The backward walks the steps in reverse with the transposed partition of the same product: CTA r holds the gate
// gradients of its own units (pointwise in dh, dc and the saved activations), multiplies that [rows x H/4] slice by
// its 128 x 512 slice of W_hh (registers again) into a PARTIAL dh_{t-1}[rows, H], and sends every 32-unit column block
// to the CTA that owns those units (st.async, fp32); the owner adds the 16 partial blocks in rank order.  dW_hh, dW_ih,
// db and dX are GEMMs / column sums over the saved dG ([T*B, 4H] bf16) afterwards.
// Global traffic stays off the step's critical path: the next step's inputs arrive by cp.async while the current one
// computes, results are staged in shared memory and written out with 16-byte stores after the sends.
// Sequence-major buffers ([T][B][...]) everywhere inside; the output is written batch-first for the encoder.
#include "tc_common.cuh"
#include "../../include/mmnas_b200.h"

namespace {

constexpr int CL = 16;                // CTAs per cluster (non-portable size, opt-in)
constexpr int THREADS = 256;          // 8 warps

struct LstmArgs {
  int T, B, H, R;                // R = batch rows per cluster (16 or 32)
  const float* xw;               // [T*B, 4H]  input projection + both biases (sequence-major rows t*B + b)
  const __nv_bfloat16* whh;      // W_hh [4H, H]
  float* out;                    // [B, T, H]  fp32, batch-first
  __nv_bfloat16* out16;          // [B*T, H]   bf16 shadow of out (or null)
  __nv_bfloat16* h16;            // [(T+1)*B, H]  slice 0 = zeros, slice t+1 = h_t
  float* gates;                  // [T*B, 4H]  activated i, f, g, o
  float* cell;                   // [T*B, H]   c_t
  const float* dout;             // [B, T, H]
  __nv_bfloat16* dG;             // [T*B, 4H]  gate pre-activation gradients
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// tanh.approx.f32: one MUFU op, max relative error 2^-11 — below the bf16 rounding of h and far inside the arm's 2e-2.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// remote (or local) shared-memory store that completes its byte count on the destination CTA's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t addr, float x, float y, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
               ::"r"(addr), "r"(__float_as_uint(x)), "r"(__float_as_uint(y)), "r"(mbar) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// mbarrier wait with a bound: a protocol error must end as a wrong answer in a test, never as a hung GPU
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (int i = 0; i < (1 << 16) && !done; ++i)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// shared rows of KD bf16 with a pitch of KD * 2 + 64 bytes: the eight lanes of a quarter-warp (rows {g, g'} x k-offsets
// {0, 16, 32, 48} bytes) hit eight different 16-byte bank groups
__host__ __device__ __forceinline__ uint32_t row_pitch(int kd) { return (uint32_t)kd * 2u + 64u; }

// ------------------------------------------------------------------------------------------------ forward
// NT = n-tiles (unit pairs) per warp; H = 256 * NT; units per CTA = 16 * NT.
template <int NT>
__global__ void __launch_bounds__(THREADS, 1) lstm_fwd_kernel(LstmArgs a) {
  constexpr int H = 256 * NT, UPC = 16 * NT, KC = H / 32;
  // fp32 staging rows of UPC values, pitch UPC + 4 words: 16-byte aligned for the vector flush / cp.async, and the 32
  // lanes of a fragment access (8 rows x 4 gates, one unit) spread over 8 banks instead of 1
  constexpr int PG = UPC + 4, S4 = UPC / 4, SEG = UPC / 8;
  extern __shared__ __align__(16) uint8_t smem[];
  const int T = a.T, B = a.B, R = a.R, MT = R >> 4;
  const uint32_t rank = cluster_ctarank();
  const int row0 = (blockIdx.x / CL) * R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const uint32_t pitch = row_pitch(H);
  const uint32_t hbuf = smem_u32(smem);                          // [2][R] rows of h (bf16)
  uint8_t* hs = smem + 2u * R * pitch;                           // [R][UPC] bf16: this CTA's new h slice
  float* sxw = reinterpret_cast<float*>(hs + (size_t)R * UPC * 2);            // [R][4] rows: xw of the current step
  float* sg = sxw + (size_t)R * 4 * PG;                          // [R][4] rows: activated gates
  float* sc = sg + (size_t)R * 4 * PG;                           // [R] rows: c_t
  float* sh = sc + (size_t)R * PG;                               // [R] rows: h_t (fp32)
  const uint32_t bars = smem_u32(sh + (size_t)R * PG);           // 2 mbarriers: h buffer b complete
  const int u0 = (int)rank * UPC;

  auto fetch_xw = [&](int t) {                                   // xw_t[rows of the cluster][4 gates][own units] -> sxw
    for (int i = tid; i < R * 4 * S4; i += THREADS) {
      const int rq = i / S4, s4 = i % S4, b = min(row0 + (rq >> 2), B - 1);
      cp_async16(smem_u32(sxw + (size_t)rq * PG + 4 * s4), a.xw + ((size_t)t * B + b) * (4 * H) + (rq & 3) * H + u0 + 4 * s4);
    }
    cp_async_commit();
  };
  fetch_xw(0);
  if (tid == 0) {
    mbar_init(bars, 1); mbar_init(bars + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // W_hh fragments: n-tile j of this warp = unit pair p = warp * NT + j; fragment column g -> gate g >> 1, unit parity g & 1
  uint4 wreg[NT][KC];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int unit = u0 + 2 * (warp * NT + j) + (g & 1), gate = g >> 1;
    const __nv_bfloat16* wrow = a.whh + (size_t)(gate * H + unit) * H + 8 * t4;
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) wreg[j][kc] = __ldg(reinterpret_cast<const uint4*>(wrow + 32 * kc));
  }
  // slice 0 of h16 (h_{-1} = 0) for this CTA's rows and units: the weight-gradient GEMM of the backward reads it
  for (int i = tid; i < R * SEG; i += THREADS) {
    const int r = row0 + i / SEG;
    if (r < B) *reinterpret_cast<uint4*>(a.h16 + (size_t)r * H + u0 + 8 * (i % SEG)) = make_uint4(0u, 0u, 0u, 0u);
  }
  float cst[2][NT];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j) cst[m][j] = 0.f;
  cl_sync();                                                     // every CTA runs and has initialised its mbarriers

  for (int t = 0; t < T; ++t) {
    const uint32_t hcur = hbuf + (uint32_t)(t & 1) * R * pitch, hnxt = hbuf + (uint32_t)((t + 1) & 1) * R * pitch;
    const bool last = t + 1 == T;
    if (!last && tid == 0) mbar_expect_tx(bars + 8u * ((t + 1) & 1), (uint32_t)R * H * 2u);      // h_t: R x H bf16 from 16 CTAs
    float acc[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[m][j][0] = acc[m][j][1] = acc[m][j][2] = acc[m][j][3] = 0.f;
    if (t > 0) {
      mbar_wait_bounded(bars + 8u * (t & 1), (uint32_t)((t - 1) >> 1) & 1u);       // the 16 slices of h_{t-1} have landed
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (m >= MT) break;
        const uint32_t pa = hcur + (uint32_t)(16 * m + g) * pitch + (uint32_t)t4 * 16u, pb = pa + 8u * pitch;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          const uint4 va = lds_v4(pa + 64u * kc), vb = lds_v4(pb + 64u * kc);
          const uint32_t a1[4] = {va.x, vb.x, va.y, vb.y};
          const uint32_t a2[4] = {va.z, vb.z, va.w, vb.w};
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            mma16816(acc[m][j], a1, wreg[j][kc].x, wreg[j][kc].y);
            mma16816(acc[m][j], a2, wreg[j][kc].z, wreg[j][kc].w);
          }
        }
      }
    }
    cp_async_wait_all();
    __syncthreads();                                             // xw_t staged; last step's flush has left sg / sc / sh / hs
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      if (m >= MT) break;
      const int l0 = 16 * m + g, l1 = l0 + 8;                    // local rows of this lane's fragment
      // this lane: gate t4 of units (2p, 2p + 1) for rows l0 (acc[.][0..1]) and l1 (acc[.][2..3])
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int p = warp * NT + j;
        float act[4];
        const float ka = (t4 == 2) ? 1.f : 0.5f, kb = (t4 == 2) ? 0.f : 0.5f;     // tanh(v)  or  sigmoid(v) = 0.5 tanh(v / 2) + 0.5
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int lr = (e < 2) ? l0 : l1, ul = 2 * p + (e & 1);
          const float v = acc[m][j][e] + sxw[((size_t)lr * 4 + t4) * PG + ul];
          act[e] = fmaf(tanh_fast(v * ka), ka, kb);
          sg[((size_t)lr * 4 + t4) * PG + ul] = act[e];
        }
        // lane t4 of the quad owns combination e = t4: (row, unit) = (e < 2 ? l0 : l1, 2p + (e & 1)); gather its gates
        float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f;
        const int qb = lane & ~3;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float vi = __shfl_sync(0xffffffffu, act[e], qb + 0);
          const float vf = __shfl_sync(0xffffffffu, act[e], qb + 1);
          const float vg = __shfl_sync(0xffffffffu, act[e], qb + 2);
          const float vo = __shfl_sync(0xffffffffu, act[e], qb + 3);
          if (e == t4) { gi = vi; gf = vf; gg = vg; go = vo; }
        }
        const int lrow = (t4 < 2) ? l0 : l1, ul = 2 * p + (t4 & 1);
        const float cn = gf * cst[m][j] + gi * gg;
        cst[m][j] = cn;
        const float hn = go * tanh_fast(cn);
        *reinterpret_cast<__nv_bfloat16*>(hs + ((uint32_t)lrow * UPC + ul) * 2u) = __float2bfloat16_rn(hn);
        sc[lrow * PG + ul] = cn;
        sh[lrow * PG + ul] = hn;
      }
    }
    __syncthreads();                                             // the CTA's slice of step t is complete in shared memory
    if (!last) {
      fetch_xw(t + 1);
      // the slice [R][UPC] bf16 into the next-step buffer of all 16 CTAs, 16 bytes per store; every store counts on the
      // receiver's barrier
      const uint32_t rbar = bars + 8u * ((t + 1) & 1);
      for (int i = tid; i < CL * R * SEG; i += THREADS) {
        const uint32_t dest = (uint32_t)(i / (R * SEG));
        const int v = i % (R * SEG), lr = v / SEG, sgm = v % SEG;
        const uint4 val = *reinterpret_cast<const uint4*>(hs + ((uint32_t)lr * UPC + 8 * sgm) * 2u);
        st_async_v4(mapa_shared(hnxt + (uint32_t)lr * pitch + (uint32_t)(u0 + 8 * sgm) * 2u, dest), val, mapa_shared(rbar, dest));
      }
    }
    // flush the step to global memory (saved activations for the backward, the output sequence, h_t as bf16)
    for (int i = tid; i < R * 4 * S4; i += THREADS) {
      const int rq = i / S4, s4 = i % S4, b = row0 + (rq >> 2);
      if (b < B)
        *reinterpret_cast<float4*>(a.gates + ((size_t)t * B + b) * (4 * H) + (rq & 3) * H + u0 + 4 * s4) =
            *reinterpret_cast<const float4*>(sg + (size_t)rq * PG + 4 * s4);
    }
    for (int i = tid; i < R * S4; i += THREADS) {
      const int lr = i / S4, s4 = i % S4, b = row0 + lr;
      if (b < B) {
        *reinterpret_cast<float4*>(a.cell + ((size_t)t * B + b) * H + u0 + 4 * s4) = *reinterpret_cast<const float4*>(sc + lr * PG + 4 * s4);
        *reinterpret_cast<float4*>(a.out + ((size_t)b * T + t) * H + u0 + 4 * s4) = *reinterpret_cast<const float4*>(sh + lr * PG + 4 * s4);
      }
    }
    for (int i = tid; i < R * SEG; i += THREADS) {
      const int lr = i / SEG, sgm = i % SEG, b = row0 + lr;
      if (b < B) {
        const uint4 val = *reinterpret_cast<const uint4*>(hs + ((uint32_t)lr * UPC + 8 * sgm) * 2u);
        *reinterpret_cast<uint4*>(a.h16 + ((size_t)(t + 1) * B + b) * H + u0 + 8 * sgm) = val;
        if (a.out16) *reinterpret_cast<uint4*>(a.out16 + ((size_t)b * T + t) * H + u0 + 8 * sgm) = val;
      }
    }
  }
  cl_sync();                                                     // no CTA exits while a peer may still address its memory
}

// ------------------------------------------------------------------------------------------------ backward
// CTA r: dG_t[rows, own gate columns] (local) x W_hh[own gate columns, all H units] -> partial dh_{t-1}[rows, H]; warp w
// covers units [w H/8, (w+1) H/8) as NTW n-tiles; K = 4 UPC own gate columns (local index q * UPC + unit).
template <int NT>
__global__ void __launch_bounds__(THREADS, 1) lstm_bwd_kernel(LstmArgs a) {
  constexpr int H = 256 * NT, UPC = 16 * NT, K4 = 4 * H, KL = 4 * UPC, KCB = KL / 32, NTW = H / 64;
  constexpr int PP = UPC / 8;                                    // (row, unit) pairs per thread at R = 32: R * UPC / 256
  constexpr int S4 = UPC / 4, SEG = UPC / 8;
  extern __shared__ __align__(16) uint8_t smem[];
  const int T = a.T, B = a.B, R = a.R, MT = R >> 4;
  const uint32_t rank = cluster_ctarank();
  const int row0 = (blockIdx.x / CL) * R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const uint32_t pitch = row_pitch(KL);
  const uint32_t dgs = smem_u32(smem);                           // [R] rows: this CTA's dG slice (bf16, local k = q UPC + unit)
  float* red = reinterpret_cast<float*>(smem + (size_t)R * pitch);            // [2][16 sources][R][UPC] partial dh (fp32)
  float* sgt = red + 2 * (size_t)CL * R * UPC;                   // [R][4][UPC] saved gates of the step
  float* sct = sgt + (size_t)R * 4 * UPC;                        // [R][UPC] c_t
  float* scp = sct + (size_t)R * UPC;                            // [R][UPC] c_{t-1}
  float* sdo = scp + (size_t)R * UPC;                            // [R][UPC] dout[:, t]
  const uint32_t bars = smem_u32(sdo + (size_t)R * UPC);         // 2 mbarriers: partial-sum buffer b complete
  const int u0 = (int)rank * UPC;

  auto fetch = [&](int t) {
    for (int i = tid; i < R * 4 * S4; i += THREADS) {
      const int rq = i / S4, s4 = i % S4, b = min(row0 + (rq >> 2), B - 1);
      cp_async16(smem_u32(sgt + (size_t)rq * UPC + 4 * s4), a.gates + ((size_t)t * B + b) * K4 + (rq & 3) * H + u0 + 4 * s4);
    }
    for (int i = tid; i < R * S4; i += THREADS) {
      const int lr = i / S4, s4 = i % S4, b = min(row0 + lr, B - 1);
      cp_async16(smem_u32(sct + lr * UPC + 4 * s4), a.cell + ((size_t)t * B + b) * H + u0 + 4 * s4);
      if (t > 0) cp_async16(smem_u32(scp + lr * UPC + 4 * s4), a.cell + ((size_t)(t - 1) * B + b) * H + u0 + 4 * s4);
      cp_async16(smem_u32(sdo + lr * UPC + 4 * s4), a.dout + ((size_t)b * T + t) * H + u0 + 4 * s4);
    }
    cp_async_commit();
  };
  fetch(T - 1);
  if (tid == 0) {
    mbar_init(bars, 1); mbar_init(bars + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // weight fragments straight from W_hh [4H, H]: n-tile j of this warp = units n0 .. n0 + 7, n0 = warp H/8 + 8j (column g =
  // unit n0 + g); k chunk c covers local k = 32c .. 32c + 31, i.e. gate q = 32c / UPC, own units (32c % UPC) .. — eight
  // 2-byte loads a fragment register pair, once per launch (1 MB of weights, L2-resident)
  uint4 wreg[NTW][KCB];
#pragma unroll
  for (int j = 0; j < NTW; ++j) {
    const unsigned short* wcol = reinterpret_cast<const unsigned short*>(a.whh) + warp * (H / 8) + 8 * j + g;
#pragma unroll
    for (int c = 0; c < KCB; ++c) {
      const int kl = 32 * c + 8 * t4, q = kl / UPC, ul = kl % UPC;
      const unsigned short* w0 = wcol + (size_t)(q * H + u0 + ul) * H;
      uint32_t e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) e[i] = __ldg(w0 + (size_t)i * H);
      wreg[j][c] = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
    }
  }
  float dcr[PP];
#pragma unroll
  for (int i = 0; i < PP; ++i) dcr[i] = 0.f;
  const int npairs = R * UPC;                                    // pair p = tid + 256 i  ->  local row p / UPC, unit p % UPC
  cl_sync();                                                     // every CTA runs and has initialised its mbarriers

  for (int t = T - 1; t >= 0; --t) {
    const float* rbuf = red + (size_t)(t & 1) * CL * R * UPC;
    cp_async_wait_all();
    if (t < T - 1) mbar_wait_bounded(bars + 8u * (t & 1), (uint32_t)((T - 2 - t) >> 1) & 1u);     // 16 partial blocks of dh_t landed
    __syncthreads();                                             // inputs of step t staged; the last product has left dgs
    // gate gradients of the own units (pointwise)
#pragma unroll
    for (int i = 0; i < PP; ++i) {
      const int p = tid + THREADS * i;
      if (p >= npairs) break;
      const int lr = p / UPC, ul = p % UPC;
      const float gi = sgt[((size_t)lr * 4 + 0) * UPC + ul], gf = sgt[((size_t)lr * 4 + 1) * UPC + ul];
      const float gg = sgt[((size_t)lr * 4 + 2) * UPC + ul], go = sgt[((size_t)lr * 4 + 3) * UPC + ul];
      const float ct = sct[p];
      const float cp = t > 0 ? scp[p] : 0.f;
      float dh = sdo[p];
      if (t < T - 1) {
#pragma unroll
        for (int src = 0; src < CL; ++src) dh += rbuf[(size_t)src * R * UPC + p];      // fixed order: bit-reproducible
      }
      const float tc = tanh_fast(ct);
      const float dc = dh * go * (1.f - tc * tc) + dcr[i];
      dcr[i] = dc * gf;
      uint8_t* sd = smem + (uint32_t)lr * pitch + (uint32_t)ul * 2u;
      *reinterpret_cast<__nv_bfloat16*>(sd) = __float2bfloat16_rn(dc * gg * gi * (1.f - gi));
      *reinterpret_cast<__nv_bfloat16*>(sd + UPC * 2) = __float2bfloat16_rn(dc * cp * gf * (1.f - gf));
      *reinterpret_cast<__nv_bfloat16*>(sd + 2 * UPC * 2) = __float2bfloat16_rn(dc * gi * (1.f - gg * gg));
      *reinterpret_cast<__nv_bfloat16*>(sd + 3 * UPC * 2) = __float2bfloat16_rn(dh * tc * go * (1.f - go));
    }
    __syncthreads();                                             // slice complete; every read of the step's inputs is done
    if (t > 0) {
      fetch(t - 1);
      if (tid == 0) mbar_expect_tx(bars + 8u * ((t - 1) & 1), (uint32_t)CL * R * UPC * 4u);
      const uint32_t rnext = smem_u32(red + (size_t)((t - 1) & 1) * CL * R * UPC) + (uint32_t)rank * R * UPC * 4u;
      const uint32_t rbar = bars + 8u * ((t - 1) & 1);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (m >= MT) break;
        const int l0 = 16 * m + g, l1 = l0 + 8;
        float acc[NTW][4];
#pragma unroll
        for (int j = 0; j < NTW; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        const uint32_t pa = dgs + (uint32_t)l0 * pitch + (uint32_t)t4 * 16u, pb = pa + 8u * pitch;
#pragma unroll
        for (int c = 0; c < KCB; ++c) {
          const uint4 va = lds_v4(pa + 64u * c), vb = lds_v4(pb + 64u * c);
          const uint32_t a1[4] = {va.x, vb.x, va.y, vb.y};
          const uint32_t a2[4] = {va.z, vb.z, va.w, vb.w};
#pragma unroll
          for (int j = 0; j < NTW; ++j) {
            mma16816(acc[j], a1, wreg[j][c].x, wreg[j][c].y);
            mma16816(acc[j], a2, wreg[j][c].z, wreg[j][c].w);
          }
        }
        // C fragment: rows l0 / l1, units n0 + 2 t4, + 1 -> the owner of those units, block [source = this rank]
#pragma unroll
        for (int j = 0; j < NTW; ++j) {
          const int unit = warp * (H / 8) + 8 * j + 2 * t4;
          const uint32_t dest = (uint32_t)(unit / UPC), ul = (uint32_t)(unit % UPC);
          const uint32_t base = mapa_shared(rnext, dest), dbar = mapa_shared(rbar, dest);
          st_async_v2(base + ((uint32_t)l0 * UPC + ul) * 4u, acc[j][0], acc[j][1], dbar);
          st_async_v2(base + ((uint32_t)l1 * UPC + ul) * 4u, acc[j][2], acc[j][3], dbar);
        }
      }
    }
    // dG_t to global memory (the weight / input gradient GEMMs read it afterwards)
    for (int i = tid; i < R * 4 * SEG; i += THREADS) {
      const int rq = i / SEG, sgm = i % SEG, lr = rq >> 2, q = rq & 3, b = row0 + lr;
      if (b < B)
        *reinterpret_cast<uint4*>(a.dG + ((size_t)t * B + b) * K4 + q * H + u0 + 8 * sgm) =
            *reinterpret_cast<const uint4*>(smem + (uint32_t)lr * pitch + (uint32_t)(q * UPC + 8 * sgm) * 2u);
    }
  }
  cl_sync();
}

int check(int T, int B, int H) {
  MMNAS_CHECK_ARG(T >= 1 && B >= 1 && B <= 256, "lstm: need 1 <= B <= 256, T >= 1");
  if (H != 256 && H != 512) {
    mmnas_set_error("lstm: H must be 256 or 512");
    return MMNAS_ERR_UNSUPPORTED;
  }
  return MMNAS_OK;
}

inline int rows_per_cluster(int B) { return B <= 128 ? 16 : 32; }

template <typename K>
int launch_cluster(K kernel, size_t smem, LstmArgs& a, cudaStream_t s) {
  MMNAS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MMNAS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL * ceil_div(a.B, a.R)); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int nclusters = 0;
  MMNAS_CUDA(cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg));
  if (nclusters < 1) {
    mmnas_set_error("lstm: a cluster of 16 CTAs cannot be scheduled on this device");
    return MMNAS_ERR_UNSUPPORTED;
  }
  mmnas_count_launch();
  MMNAS_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
  return MMNAS_OK;
}

unsigned long long al256(unsigned long long v) { return (v + 255ull) & ~255ull; }

void carve(LstmArgs& a, void* ws) {
  const unsigned long long tb = (unsigned long long)a.T * a.B;
  uint8_t* p = static_cast<uint8_t*>(ws);
  a.h16 = reinterpret_cast<__nv_bfloat16*>(p); p += al256((tb + a.B) * a.H * 2);
  a.gates = reinterpret_cast<float*>(p); p += al256(tb * 4 * a.H * 4);
  a.cell = reinterpret_cast<float*>(p); p += al256(tb * a.H * 4);
  a.dG = reinterpret_cast<__nv_bfloat16*>(p);
}

}  // namespace

extern "C" int mmnas_lstm_workspace(int T, int B, int H, unsigned long long* bytes) {
  if (int rc = check(T, B, H)) return rc;
  MMNAS_CHECK_ARG(bytes, "lstm_workspace: null pointer");
  // h16 [(T+1)*B, H] bf16 | gates [T*B, 4H] f32 | cell [T*B, H] f32 | dG [T*B, 4H] bf16
  const unsigned long long tb = (unsigned long long)T * B;
  *bytes = al256((tb + B) * H * 2) + al256(tb * 4 * H * 4) + al256(tb * H * 4) + al256(tb * 4 * H * 2);
  return MMNAS_OK;
}

extern "C" int mmnas_lstm_fwd(int T, int B, int H, const float* xw, const void* whh16, float* out, void* out16,
                              void* workspace, mmnas_stream stream) {
  if (int rc = check(T, B, H)) return rc;
  MMNAS_CHECK_ARG(xw && whh16 && out && workspace, "lstm_fwd: null pointer");
  LstmArgs a = {};
  a.T = T; a.B = B; a.H = H; a.R = rows_per_cluster(B); a.xw = xw; a.whh = static_cast<const __nv_bfloat16*>(whh16);
  a.out = out; a.out16 = static_cast<__nv_bfloat16*>(out16);
  carve(a, workspace);
  const int upc = H / CL, pg = upc + 4;
  // h double buffer | h slice (bf16) | xw, gates (fp32 [R][4] rows of upc + 4) | c, h (fp32 [R] rows) | 2 mbarriers
  const size_t smem = 2 * (size_t)a.R * row_pitch(H) + (size_t)a.R * upc * 2 + 2 * (size_t)a.R * 4 * pg * 4 + 2 * (size_t)a.R * pg * 4 + 16;
  if (H == 512) return launch_cluster(lstm_fwd_kernel<2>, smem, a, (cudaStream_t)stream);
  return launch_cluster(lstm_fwd_kernel<1>, smem, a, (cudaStream_t)stream);
}

extern "C" int mmnas_lstm_bwd(int T, int B, int H, const float* dout, const void* whh16, void* workspace,
                              mmnas_stream stream) {
  if (int rc = check(T, B, H)) return rc;
  MMNAS_CHECK_ARG(dout && whh16 && workspace, "lstm_bwd: null pointer");
  LstmArgs a = {};
  a.T = T; a.B = B; a.H = H; a.R = rows_per_cluster(B); a.dout = dout; a.whh = static_cast<const __nv_bfloat16*>(whh16);
  carve(a, workspace);
  const int upc = H / CL;
  // dG slice rows | partial dh [2][16][R][upc] fp32 | saved gates [R][4][upc] | c_t, c_{t-1}, dout [R][upc] | 2 mbarriers
  const size_t smem = (size_t)a.R * row_pitch(4 * upc) + 2 * (size_t)CL * a.R * upc * 4 + (size_t)a.R * 4 * upc * 4 + 3 * (size_t)a.R * upc * 4 + 16;
  if (H == 512) return launch_cluster(lstm_bwd_kernel<2>, smem, a, (cudaStream_t)stream);
  return launch_cluster(lstm_bwd_kernel<1>, smem, a, (cudaStream_t)stream);
}
