/* libmmnas_b200 — C ABI of the B200-native MMnas operator hot path.
 *
 * The reference (MILVLG/mmnas) has no FFI of its own: its operator path is pure PyTorch
 * (mmnas/model/modules.py, mmnas/model/mixed.py).  These entry points are what a binding for that path
 * binds instead of torch's nn.Linear / torch.matmul / F.softmax / masked_fill / nn.Dropout / x.std calls;
 * each one cites the reference lines it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - `stream` is a cudaStream_t; every call is asynchronous on it, allocates nothing, keeps no state
 *     between calls and is safe to capture in a CUDA graph;
 *   - return 0 on success, <0 on error (MMNAS_ERR_*); mmnas_last_error() gives a thread-local message;
 *   - dtype arguments: 0 = float32, 1 = bfloat16 (storage type; all arithmetic accumulates in fp32);
 *   - dropout: `rng_state` is a device array {seed, step} of two uint64 (NULL or p == 0 disables it);
 *     `salt` identifies the call site; element i of a site is kept iff hash16(seed, step, salt, i) >= p*65536
 *     and scaled by 1/(1-p).  Forward and backward of one site pass identical (rng_state, salt, p).
 *   - tensors are row-major; "ld" is a row pitch in ELEMENTS.
 */
#ifndef MMNAS_B200_H
#define MMNAS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MMNAS_B200_ABI_VERSION 8

typedef void* mmnas_stream;

int mmnas_abi_version(void);
const char* mmnas_last_error(void);
/* number of CUDA kernels this library has launched in this process so far (memsets not counted) */
unsigned long long mmnas_launch_count(void);

/* ---- dense contractions: nn.Linear forward/backward (modules.py:18,38,172-175,216-220) ----------------
 * fp32 arm (FFMA): C[M,N] = epi(A @ B (+bias[N])) (+C);  A(m,k) = A[m*a_rs + k*a_cs], B(k,n) = B[k*b_rs + n*b_cs].
 * epilogue: 0 none, 1 ReLU (FC.relu :27), 2 ReLU+dropout (:29), 3 C = aux[m,n] > 0 ? v*aux_scale : 0 (ReLU/dropout backward). */
int mmnas_gemm_f32(int M, int N, int K, const float* A, long a_rs, long a_cs, const float* B, long b_rs, long b_cs,
                   float* C, long ldc, const float* bias, int epilogue, int accumulate, const float* aux, long ld_aux,
                   float aux_scale, const unsigned long long* rng_state, unsigned long long salt, float p,
                   mmnas_stream stream);

/* bf16 arm (tcgen05/TMEM/TMA).  A: a_mn_major 0 -> stored [M][K] (pitch lda), 1 -> stored [K][M];
 * B: b_mn_major 0 -> stored [N][K] (pitch ldb), 1 -> stored [K][N].  C fp32 or bf16 [M][N] (pitch ldc).
 * relu / dropout / aux-mask (aux is bf16 [M][N]) / accumulate (fp32 C += result) as above; split_k > 1 adds
 * the partial products into a caller-zeroed fp32 C with red.add (weight gradients, reduction over tokens).
 * Requirements: N % 32 == 0, pitches % 8 == 0, 16-byte aligned bases. */
int mmnas_gemm_bf16(int M, int N, int K, const void* A, long lda, int a_mn_major, const void* B, long ldb,
                    int b_mn_major, void* C, long ldc, int out_bf16, const float* bias, int relu, int accumulate,
                    const void* aux, long ld_aux, float aux_scale, int split_k, const unsigned long long* rng_state,
                    unsigned long long salt, float p, mmnas_stream stream);

/* bf16 arm, projection with the block tail fused into its epilogue (tcgen05 cta_group::2 pairs in a cluster that owns
 * whole rows; row statistics across the cluster through distributed shared memory):
 *   z = x + dropout(A W^T + bias) ; out = gamma (z - mean) / (std_unbiased + eps) + beta        (modules.py:261-271, :52-56)
 * A [M,K] bf16 (pitch lda), W [N,K] bf16 (pitch ldb), bias [N] or NULL, x [M,N] fp32 residual input or NULL; writes z [M,N]
 * fp32 (saved for the backward), out [M,N] fp32, out_bf16 [M,N], mean / sigma [M].  N must be 256 or 512 and gamma, beta,
 * out_bf16 non-NULL; otherwise returns MMNAS_ERR_UNSUPPORTED (-2) and the caller runs mmnas_gemm_bf16 +
 * mmnas_ln_residual_fwd.  Dropout stream identical to mmnas_ln_residual_fwd's. */
int mmnas_gemm_ln_bf16(int M, int N, int K, const void* A, long lda, const void* W, long ldb, const float* bias,
                       const float* x, const float* gamma, const float* beta, float eps, float* z, float* out,
                       void* out_bf16, float* mean, float* sigma, const unsigned long long* rng_state,
                       unsigned long long salt, float p, mmnas_stream stream);

/* ---- attention core: MHAtt.att (modules.py:190-199) / RelMHAtt.forward (:232-240) ---------------------
 * q [B*Nq rows, pitch ldq], k/v [B*Nk rows]; head h reads columns [h*64, h*64+64).  kmask [B,Nk] bytes, 1 = padded
 * key (masked_fill(mask, -1e9) AFTER the bias add); bias [B,heads,Nq,Nk] fp32 or NULL; o [B*Nq rows, pitch ldo]
 * in merged-head layout.  Nk <= 128, head_dim == 64. */
int mmnas_attn_fwd(int dtype, int B, int heads, int Nq, int Nk, int head_dim, const void* q, long ldq, const void* k,
                   long ldk, const void* v, long ldv, const unsigned char* kmask, const float* bias, void* o, long ldo,
                   float scale, const unsigned long long* rng_state, unsigned long long salt, float p,
                   mmnas_stream stream);
/* Backward: recomputes the attention map; dbias (fp32 [B,heads,Nq,Nk], = dS) is written when non-NULL. */
int mmnas_attn_bwd(int dtype, int B, int heads, int Nq, int Nk, int head_dim, const void* q, long ldq, const void* k,
                   long ldk, const void* v, long ldv, const unsigned char* kmask, const float* bias, const void* o,
                   long ldo, const void* dout, long lddo, void* dq, long lddq, void* dk, long lddk, void* dv, long lddv,
                   float* dbias, float scale, const unsigned long long* rng_state, unsigned long long salt, float p,
                   mmnas_stream stream);

/* ---- RSA geometry bias: log(clamp(relu(linear_r(rel_embed)), 1e-6)) (modules.py:231,235) --------------
 * Exactly one of rel [B,N,N,R] (the reference's dense tensor) or g4 [B,N,N,4] (+Wy [R,4], by [R]: the
 * relu(linear_y_rel(.)) of full_vqa.py:103 folded in) is non-NULL.  bias out: [B,heads,N,N].  R == 64, heads <= 16.
 * mode 0: fp32 arithmetic (parity arm).  mode 1 (bf16 arm): the 64-channel contractions run on the tensor cores with
 * split-bf16 operands for r and plain bf16 operands for the gradient products (geometry input, 2-8 heads; any other
 * configuration silently uses the mode-0 kernels). */
int mmnas_relbias_fwd(int mode, int B, int N, int heads, int R, const float* rel, const float* g4, const float* Wy,
                      const float* by, const float* Wr, const float* br, float* bias, mmnas_stream stream);
/* Backward: dWr/dbr (and dWy/dby, geometry mode) are ACCUMULATED (caller zeroes); drel [B,N,N,R] written (dense mode). */
int mmnas_relbias_bwd(int mode, int B, int N, int heads, int R, const float* rel, const float* g4, const float* Wy,
                      const float* by, const float* Wr, const float* br, const float* dbias, float* drel, float* dWy,
                      float* dby, float* dWr, float* dbr, mmnas_stream stream);

/* ---- block tail: x + dropout(branch) -> LayerNorm (modules.py:261-271 and :52-56) ----------------------
 * On return `branch` holds z = x + dropout(branch) (saved for backward).  gamma/beta NULL = norm off; x NULL =
 * residual off.  out_bf16 (optional) receives a bf16 copy of `out` for the next block's tensor-core GEMM. */
int mmnas_ln_residual_fwd(int rows, int H, const float* x, float* branch, const float* gamma, const float* beta,
                          float eps, float* out, void* out_bf16, float* mean, float* sigma,
                          const unsigned long long* rng_state, unsigned long long salt, float p, mmnas_stream stream);
/* dz = grad wrt z (also the residual-path grad wrt x); dbranch = dz * dropout mask (dtype 0/1; may be NULL);
 * dgamma/dbeta are ACCUMULATED (caller zeroes).  H <= 1024. */
int mmnas_ln_residual_bwd(int rows, int H, const float* dout, const float* z, const float* mean, const float* sigma,
                          const float* gamma, float eps, float* dz, void* dbranch, int dbranch_dtype, float* dgamma,
                          float* dbeta, const unsigned long long* rng_state, unsigned long long salt, float p,
                          mmnas_stream stream);

/* ---- supernet mixed-op: MixedOp.forward 'full' mode (mixed.py:60-68) -----------------------------------
 * out = sum_k gate[k] * outs[k]   (outs: HOST array of K device pointers, each n floats; n % 4 == 0; K <= 8) */
int mmnas_mixed_accum(int K, const float* const* outs, const float* gate, float* out, long n, mmnas_stream stream);
/* gate_grad[k] = <outs[k], dout> (overwritten) — autograd's alpha_gate.grad; d_outs[k] (HOST array, entries may be
 * NULL for detached candidates) receives gate[k] * dout. */
int mmnas_mixed_alpha_dot(int K, const float* const* outs, const float* gate, const float* dout, float* gate_grad,
                          float* const* d_outs, long n, mmnas_stream stream);

/* ---- helpers -------------------------------------------------------------------------------------------- */
int mmnas_cast_f32_to_bf16(const float* src, void* dst, long n, mmnas_stream stream);
/* Batched cast: `table` is a DEVICE array of n_chunks triples of int64 {src float*, dst bf16*, count}; each chunk has
 * count <= 4096, count % 4 == 0, 16-byte aligned pointers.  Refreshes every bf16 weight shadow of a model in one launch. */
int mmnas_cast_multi(const void* table, int n_chunks, mmnas_stream stream);
/* out[c] (+)= sum_r x[r,c] (bias gradients); accumulate == 0 overwrites out, 1 adds into it (gradient buffers). */
int mmnas_colsum(int dtype, const void* x, int rows, int cols, long ld, float* out, int accumulate, mmnas_stream stream);
/* ---- stem: bf16 copy of the region features (optional, NULL to skip) + make_mask() (full_vqa.py:113-114):
 * mask[r] = 1 iff every element of row r is zero (== sum|x| == 0). */
int mmnas_cast_rowmask(const float* x, void* x_bf16, unsigned char* mask, int rows, int cols, mmnas_stream stream);
/* the mask alone, for region features that arrive as bf16 already (a loader that stores them so halves the host->device bytes) */
int mmnas_rowmask_bf16(const void* x_bf16, unsigned char* mask, int rows, int cols, mmnas_stream stream);

/* ---- geometry producer: relation_embedding (mmnas/loader/load_data_vqa.py:7-33, twins in load_data_vgd.py:7-33 and
 * load_data_itm.py:5-31) and the zero padding of load_data_vqa.py:236-239, on the device.  boxes [B,N,4] fp32
 * (x1,y1,x2,y2), pad_mask [B,N] bytes (1 = padded region; NULL = all valid) -> g4 [B,N,N,4]: log-geometry of every
 * valid pair, zeros elsewhere.  Lets a loader ship boxes instead of the N x N x 4 tensor. */
int mmnas_box_geometry(const float* boxes, const unsigned char* pad_mask, float* g4, int B, int N, mmnas_stream stream);

/* ---- text stem: the recurrence of the question / caption LSTM (full_vqa.py:68-74,94-95; full_vgd.py, full_itm.py alike:
 * nn.LSTM(WORD_EMBED_SIZE, HSIZE, num_layers=1, batch_first=True), zero initial state, output sequence only).  bf16 arm.
 * The caller computes the input projection of all steps, xw [T*B, 4H] fp32 = X W_ih^T + b_ih + b_hh with rows in
 * sequence-major order (t*B + b) and torch's gate order (i, f, g, o), e.g. with mmnas_gemm_bf16.  mmnas_lstm_fwd runs the
 * T dependent steps in ONE kernel: clusters of 16 CTAs own 16 or 32 batch rows each, W_hh [4H, H] bf16 stays in registers
 * as mma.sync fragments, h_t is exchanged through distributed shared memory (st.async on the receiver's mbarrier).  Writes
 * out [B, T, H] fp32 batch-first (+ optional bf16 copy out16 [B*T, H]).
 * The workspace (mmnas_lstm_workspace bytes, 256-byte aligned, untouched between forward and backward) keeps, in this
 * order: h16 [(T+1)*B, H] bf16 (slice 0 zero, slice t+1 = h_t), activated gates [T*B, 4H] fp32, cell states [T*B, H]
 * fp32, dG [T*B, 4H] bf16.  mmnas_lstm_bwd takes dout [B, T, H] fp32 and the same W_hh [4H, H] bf16 and
 * fills dG (gradients of the gate pre-activations, sequence-major); the caller finishes with dW_hh = dG^T h16[0:T*B],
 * dW_ih = dG^T X, db_ih = db_hh = colsum(dG), dX = dG W_ih.  1 <= B <= 256; H must be 256 or 512, otherwise
 * MMNAS_ERR_UNSUPPORTED (-2; also when the device cannot schedule a 16-CTA cluster) and the caller keeps its own LSTM. */
int mmnas_lstm_workspace(int T, int B, int H, unsigned long long* bytes);
int mmnas_lstm_fwd(int T, int B, int H, const float* xw, const void* whh_bf16, float* out, void* out_bf16, void* workspace,
                   mmnas_stream stream);
int mmnas_lstm_bwd(int T, int B, int H, const float* dout, const void* whh_bf16, void* workspace, mmnas_stream stream);

/* ---- optimizer tail: clip_grad_norm_ (train_vqa.py:310) + Adam (train_vqa.py:311 via optimizer.py:14-20) ------
 * out[0] = sum(x^2) over a flat fp32 buffer (n % 4 == 0).  Bit-reproducible (fixed summation order, no float
 * atomics), so data-parallel replicas holding the same gradients clip by the same coefficient.  `scratch` =
 * MMNAS_SUMSQ_SCRATCH floats of device memory owned by the caller, not shared between concurrent calls. */
#define MMNAS_SUMSQ_SCRATCH 1280
int mmnas_sumsq_f32(const float* x, long n, float* out, float* scratch, mmnas_stream stream);
/* One launch updates every parameter: `table` = DEVICE array of n_chunks rows of 5 int64 {param*, grad*, exp_avg*,
 * exp_avg_sq*, count} (count <= 4096).  Gradients are scaled by min(1, max_norm / (sqrt(*sumsq) + 1e-6))
 * (max_norm <= 0: no clipping); lr is read from device memory, the step count t from step_state[1] (advance it with
 * mmnas_rng_advance before the call), so the call can be replayed from a CUDA graph. */
int mmnas_clip_adam(const void* table, int n_chunks, const float* sumsq, const float* lr,
                    const unsigned long long* step_state, float beta1, float beta2, float eps, float max_norm,
                    mmnas_stream stream);
/* state[1] += 1: call once per training step so every step draws fresh dropout masks (graph-capturable). */
int mmnas_rng_advance(unsigned long long* state, mmnas_stream stream);

/* =========================================================================================================
 * Block-level entry points (ABI v7): ONE call enqueues the whole kernel sequence of a candidate block, forward or
 * backward, including the side-stream fork / join of the weight-gradient GEMMs.  They replace the forward/backward of
 *   SelfAtt.forward      modules.py:260-271     GuidedAtt.forward    modules.py:313-325   -> mmnas_mha_ln_*
 *   RelSelfAtt.forward   modules.py:286-298  (RelMHAtt :224-245)                          -> mmnas_rel_mha_ln_*
 *   FeedForward.forward  modules.py:351-362  (MLP :40-41, FC :24-31)                      -> mmnas_ffn_ln_*
 * The descriptor is a plain C struct of device pointers, sizes and scalars; the library allocates nothing: every
 * intermediate lives in a caller-provided workspace whose size mmnas_*_workspace() reports.  `workspace` is written
 * by the forward and must be handed unchanged to the backward (it holds the saved activations: fused projections,
 * attention output, z = x + dropout(branch), LayerNorm statistics, relation bias); `bwd_workspace` is scratch of the
 * backward only.  Both need 256-byte alignment.  The fork / join uses a process-wide ring of timing-disabled CUDA
 * events owned by the library (capturable in a CUDA graph).
 * ========================================================================================================= */
typedef struct mmnas_att_block {
  /* configuration */
  int precision;               /* 0 = fp32 arm (FFMA kernels), 1 = bf16 arm (tcgen05 kernels) */
  int B, Nq, Nk, H, I;         /* B samples, Nq query / Nk key tokens per sample, hidden H, inner width I = heads * 64 */
  int R;                       /* relation-embedding width (64) for RSA, 0 otherwise */
  int residual;                /* 1: out = LN(x + dropout(branch)); 0: out = LN(dropout(branch)) */
  int guided;                  /* 1: keys / values come from `kv` (GuidedAtt); 0: self-attention (Nk == Nq) */
  int accumulate_grads;        /* backward: 1 = weight / LN / linear_r gradient buffers hold running sums (+=), 0 = overwrite */
  int accumulate_geometry;     /* backward: same for dWy / dby (linear_y_rel is shared by every RSA block of a net) */
  int accumulate_dkv;          /* backward, guided: 1 = dkv += (every GuidedAtt block of a decoder adds into ONE encoder-output gradient) */
  float eps;                   /* LayerNorm eps (added to sigma) */
  float p_att, p_out;          /* dropout probability of the attention map / of the block output; 0 = off */
  unsigned long long salt_att, salt_out;
  const unsigned long long* rng_state;   /* device {seed, step}; NULL = dropout off */
  /* inputs */
  const float* x;              /* [B*Nq, H] fp32 */
  const void* x16;             /* bf16 copy of x, or NULL (bf16 arm: cast into the workspace) */
  const float* kv;             /* guided: [B*Nk, H] fp32 */
  const void* kv16;
  const unsigned char* kmask;  /* [B, Nk] bytes, 1 = padded key; NULL = no mask */
  /* parameters: fp32 masters (always) and, in the bf16 arm, bf16 copies stacked per fused GEMM */
  const float *Wq, *Wk, *Wv, *Wm;        /* [I,H] x3, [H,I] */
  const void* w16_a;           /* self: [Wv;Wk;Wq] as [3I,H];  guided: Wq [I,H] */
  const void* w16_b;           /* guided: [Wv;Wk] as [2I,H];   self: unused */
  const void* w16_m;           /* [H,I] */
  const float *ln_a, *ln_b;    /* LayerNorm a_2 / b_2, both NULL = norm off */
  const float *rel;            /* RSA, dense mode: rel_embed [B,Nq,Nq,R] */
  const float *g4, *Wy, *by;   /* RSA, geometry mode: [B,Nq,Nq,4] + linear_y_rel [R,4], [R] */
  const float *Wr, *br;        /* RSA: linear_r [heads,R], [heads] */
  /* outputs */
  float* out;                  /* [B*Nq, H] fp32 */
  void* out16;                 /* optional bf16 copy of out (next block's GEMM operand) */
  void* workspace;
  /* backward */
  const float* dout;           /* [B*Nq, H] */
  float* dx;                   /* [B*Nq, H], written */
  float* dkv;                  /* guided: [B*Nk, H], written */
  float *dWq, *dWk, *dWv, *dWm, *dln_a, *dln_b, *dWy, *dby, *dWr, *dbr;
  float* drel;                 /* dense mode: [B,Nq,Nq,R], written */
  void* bwd_workspace;
  mmnas_stream stream;
  mmnas_stream side_stream;    /* bf16-arm weight-gradient GEMMs run here, concurrently; NULL = everything on `stream` */
} mmnas_att_block;

typedef struct mmnas_ffn_block {
  int precision;
  int M, H, F;                 /* M tokens, hidden H, mid size F */
  int residual;
  int accumulate_grads;
  float eps;
  float p_mid, p_out;          /* dropout of the hidden activation (FC :29) / of the block output (:353) */
  unsigned long long salt_mid, salt_out;
  const unsigned long long* rng_state;
  const float* x; const void* x16;
  const float *W1, *b1, *W2, *b2;        /* mlp.fc.linear [F,H],[F]; mlp.linear [H,F],[H] */
  const void *w16_1, *w16_2;
  const float *ln_a, *ln_b;
  float* out; void* out16;
  void* workspace;
  const float* dout; float* dx;
  float *dW1, *db1, *dW2, *db2, *dln_a, *dln_b;
  void* bwd_workspace;
  mmnas_stream stream, side_stream;
} mmnas_ffn_block;

/* sizeof() of the two descriptors as compiled into the library (bindings check their mirror structs against it) */
int mmnas_att_block_sizeof(void);
int mmnas_ffn_block_sizeof(void);
/* bytes of `workspace` / `bwd_workspace` for this configuration (only the configuration fields are read) */
int mmnas_att_block_workspace(const mmnas_att_block* d, unsigned long long* fwd_bytes, unsigned long long* bwd_bytes);
int mmnas_ffn_block_workspace(const mmnas_ffn_block* d, unsigned long long* fwd_bytes, unsigned long long* bwd_bytes);
/* SelfAtt / GuidedAtt (d->R must be 0) */
int mmnas_mha_ln_fwd(const mmnas_att_block* d);
int mmnas_mha_ln_bwd(const mmnas_att_block* d);
/* RelSelfAtt (d->R == 64, exactly one of rel / g4 given) */
int mmnas_rel_mha_ln_fwd(const mmnas_att_block* d);
int mmnas_rel_mha_ln_bwd(const mmnas_att_block* d);
/* FeedForward */
int mmnas_ffn_ln_fwd(const mmnas_ffn_block* d);
int mmnas_ffn_ln_bwd(const mmnas_ffn_block* d);

#ifdef __cplusplus
}
#endif
#endif
