/* mclip.h — C ABI of the B200-native Mammo-CLIP contrastive pre-training hot path.
 *
 * The reference (batmanlab/Mammo-CLIP, pure Python over PyTorch) has no FFI of its own; its plug-in surface for this
 * path is the set of Python factories build_model / load_image_encoder / load_text_encoder / load_projection_head /
 * build_loss (breastclip/model/__init__.py:10, model/modules/__init__.py:11,59,78, loss/__init__.py:9).  This header
 * is the boundary we define UNDER those factories: plain pointers and sizes, no torch types.  Each entry point cites
 * the reference code it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory (incl. workspaces)
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and allocate nothing
 *   - return value: 0 on success, <0 on error (mclip_last_error() has the text); there is no CPU fallback
 *   - activations are NHWC bf16, statistics / embeddings / losses / master weights fp32
 */
#ifndef MCLIP_H_
#define MCLIP_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------------------------- */
const char* mclip_last_error(void);           /* text of the last error on this thread */
int mclip_version(void);                      /* ABI version, bumped on any signature change */
int mclip_device_check(void);                 /* 0 iff the current device is sm_100 (B200) */

/* ---- fused InfoNCE (+ NVLink all-gather) ----------------------------------------------------------------------
 * Replaces BreastClip_contrastive.forward (loss/breast_clip_contrastive.py:28-59), BreastClip.forward
 * (loss/breast_clip.py:29-127) and DistAutogradAllGatherFunction (util/dist_autograd.py:4-26): forward AND backward
 * in one cooperative kernel.  A pair (a,b) scores tensor a's rows against tensor b's rows over all world*batch
 * samples; w_row weighs the row-wise CE (reference: `scale * x_a @ all_b.T`), w_col the column-wise CE
 * (`scale * x_b @ all_a.T`).  out = [loss, dloss/dscale, then per pair (row CE, col CE)].  grad[k] receives
 * d(sum over ranks of loss)/d local[k], which is what the reference's reduce_scatter(SUM) delivers. */
#define MCLIP_LOSS_MAX_TENSORS 4
#define MCLIP_LOSS_MAX_PAIRS 8
#define MCLIP_LOSS_MAX_WORLD 16
typedef struct mclip_loss_args {
  int world, rank, batch, dim;                 /* W, r, B (per rank), D */
  int n_tensors, n_pairs;
  const float* local[MCLIP_LOSS_MAX_TENSORS];  /* [B,D] fp32, this rank's (L2-normalised) embeddings */
  float* grad[MCLIP_LOSS_MAX_TENSORS];         /* [B,D] fp32 out */
  int pair_a[MCLIP_LOSS_MAX_PAIRS], pair_b[MCLIP_LOSS_MAX_PAIRS];
  float w_row[MCLIP_LOSS_MAX_PAIRS], w_col[MCLIP_LOSS_MAX_PAIRS], label_smoothing[MCLIP_LOSS_MAX_PAIRS];
  float logit_scale;                           /* exp(logit_scale parameter), clip.py:100 */
  const float* logit_scale_dev;                /* optional DEVICE scalar, used instead of logit_scale (no host sync) */
  void* workspace; long long workspace_bytes;  /* >= mclip_loss_workspace_bytes() */
  float* out;                                  /* [2 + 2*n_pairs] fp32 */
  /* world > 1 only: symmetric (peer-mapped) gather buffers, see mclip_symm_* below */
  float* gathered[MCLIP_LOSS_MAX_TENSORS];     /* this rank's [W*B,D] gather buffer per tensor */
  float* const* peer_gathered;                 /* device table [W][n_tensors]: every rank's gather buffers */
  uint32_t* const* peer_flags;                 /* device table [W]: every rank's arrival counters (u32[W]) */
  const uint32_t* my_flags;                    /* this rank's arrival counters */
  long long epoch;                             /* 1,2,3,... identical on all ranks, +1 per call on this buffer set */
  /* ABI 4: a peer that never arrives no longer traps the context.  After peer_timeout_s seconds (0: MCLIP_PEER_TIMEOUT_S or
   * 600 s, the order of NCCL's watchdog) the kernel stores 1 + <missing rank> in *status (device int, may be NULL), writes NaN
   * to out[0] and finishes; the host raises from the status word (mammoclip_b200.ops checks it on the next call). */
  int* status; double peer_timeout_s;
  /* ABI 5: LSE exchange for large global batches (world > 1).  lse_all = this rank's [world*batch][2*MCLIP_LOSS_MAX_PAIRS] fp32
   * table inside the symmetric allocation, peer_lse = device table [W] of every rank's lse_all.  With both set, a rank scores
   * only its own rows x all columns and all rows x its own columns, pushes its 2*P*B log-sum-exps to the peers (a second arrival
   * on the same counters, so the whole job must use one mode per buffer set) and reads the others' from the table; NULL = every
   * rank evaluates the full (W*B)^2 score matrix (ABI 4 behaviour). */
  float* lse_all; float* const* peer_lse;
} mclip_loss_args;
long long mclip_loss_workspace_bytes(int world, int batch, int dim, int n_pairs);
int mclip_loss_grid(int world, int batch, int n_pairs);
int mclip_contrastive_loss(const mclip_loss_args* args, void* stream);
/* Measurement hook (bench.py --workload loss-sweep): out2 (host, may be NULL) receives {earliest push start, arrival of the last
   remote slab as first observed on this rank} in ns of %globaltimer for the call(s) since the last reset; then resets and switches
   the recording on/off. */
int mclip_loss_window(unsigned long long* out2, int enable);

/* Peer-mapped buffers for world > 1 (cudaMalloc + CUDA IPC; handles are 64 bytes, exchanged by the host). */
int mclip_ipc_alloc(long long bytes, void** out);          /* zero-filled */
int mclip_ipc_free(void* p);
int mclip_ipc_get_handle(void* p, void* handle64);
int mclip_ipc_open_handle(const void* handle64, void** out);
int mclip_ipc_close_handle(void* p);

/* ---- tcgen05 GEMMs ---------------------------------------------------------------------------------------------
 * mclip_gemm_tn: D[b,m,n] = epi(sum_k A[b,m,k] B[b|0,n,k]), bf16 in/out, fp32 accumulate in TMEM.  This is the 1x1
 * "pointwise" convolution of MBConv in NHWC (efficientnet_custom.py:105 _expand_conv, :122 _project_conv, :283
 * _conv_head; m = pixel, k = Cin, n = Cout), its data gradient (B = transposed weight), and every nn.Linear of the
 * text tower / projection heads (text_encoder.py:48 -> transformers BertModel; projection.py:28).
 * epi: + bias[n]; (optional bf16 copy of the pre-activation); act 1 = erf-GELU; + residual[b,m,n]; optional per-column (sum, sum of squares) partials of the
 * bf16-rounded output for train-mode BatchNorm (efficientnet_custom.py:106,123): stats[stat_slots][2][n]. */
typedef struct mclip_gemm_args {
  const void* a; long long lda, a_batch_stride;   /* bf16 [batches, m, k], row stride lda (elements) */
  const void* b; long long ldb, b_batch_stride;   /* bf16 [1|batches, n, k]; b_batch_stride == 0: shared by all batches */
  void* d; long long ldd, d_batch_stride;         /* bf16 [batches, m, n] */
  int m, n, k, batches;
  const float* bias;                              /* fp32 [n] or NULL */
  const void* residual; long long ldr, r_batch_stride;   /* bf16 [batches, m, n] or NULL */
  int act;                                        /* 0 none, 1 erf-GELU */
  float* stats; int stat_slots;                   /* NULL, or fp32 [stat_slots][2][n] with stat_slots from below */
  const void* dropmask; float drop_scale;         /* uint8 keep-mask [batches*m, n] applied before the residual, or NULL */
  void* aux_pre; long long ld_aux;                /* NULL, or bf16 [batches*m, ld_aux]: the value BEFORE `act` (after the bias), kept for the
                                                     activation's backward (BertIntermediate's GELU) */
  /* ABI 4: optional SECOND A operand concatenated along K: A = [a (k columns) | a2 (k2 columns)], same rows / batches.
   * b then holds [n, ceil64(k) + k2] with zeros in columns [k, ceil64(k)).  Used by the folded BatchNorm backward of the
   * expand convolution: dX = [dV0 | X] [diag(a)We ; G] (efficientnet_custom.py:105-106 autograd). */
  const void* a2; long long lda2, a2_batch_stride; int k2;
} mclip_gemm_args;
int mclip_gemm_tn_stat_slots(int m, int n, int batches);
int mclip_gemm_tn(const mclip_gemm_args* args, void* stream);

/* mclip_gemm_wgrad: out[i,j] (+)= sum_r A[r,i] B[r,j]  (bf16 in, fp32 out) — weight gradient of a 1x1 convolution /
 * Linear layer (what autograd computes for efficientnet_custom.py:105,122,283 and the BERT Linears): A = dY [pixels,
 * Cout], B = X [pixels, Cin].  Split over r across CTAs into fp32 partials, reduced in fixed order (deterministic). */
typedef struct mclip_wgrad_args {
  const void* a; long long lda;                   /* bf16 [r, i] */
  const void* b; long long ldb;                   /* bf16 [r, j] */
  float* out; long long ldo; int accumulate;      /* fp32 [i, j]; accumulate != 0: out += */
  int r, i, j;
  void* workspace; long long workspace_bytes;     /* >= mclip_gemm_wgrad_workspace_bytes() */
} mclip_wgrad_args;
long long mclip_gemm_wgrad_workspace_bytes(int r, int i, int j);
int mclip_gemm_wgrad(const mclip_wgrad_args* args, void* stream);

/* ---- depthwise / stem convolutions ------------------------------------------------------------------------------
 * mclip_dwconv_*: MBConvBlock._depthwise_conv (efficientnet_custom.py:66-73,109), a Conv2dStaticSamePadding
 * (efficient_net_custom_utils.py:248-276) with groups == channels, k in {3,5}, stride in {1,2} and the STATIC
 * (left,right,top,bottom) pads frozen at construction.  NHWC bf16.  When in_scale != NULL the input is the
 * producer's pre-BatchNorm output and swish(in_scale*y + in_shift) is applied while loading (BN :106 + swish :107
 * of the expand phase, or the stem's :273), zero padding after it.  stats: per-channel (sum, sum sq) partials of
 * the output for the following BatchNorm (:110).  Backward yields, in one pass over (dy, in): dx = gradient
 * w.r.t. the pre-activation input (already multiplied by swish'), the weight gradient, and the input BN's
 * backward reduction partials (sum dv, sum dv*yhat). */
typedef struct mclip_dwconv_args {
  int n, h, w, c, ho, wo, k, stride;
  int pad_left, pad_right, pad_top, pad_bottom;
  const void* in;                 /* bf16 [n,h,w,c] */
  const float* in_scale;          /* fp32 [c] or NULL */
  const float* in_shift;
  int in_act;                     /* 1: swish after the affine */
  const float* weight;            /* fp32 [c,1,k,k] */
  void* out;                      /* bf16 [n,ho,wo,c]                      (forward) */
  float* stats; int stat_slots;   /* fp32 [stat_slots][2][c] or NULL; slots from mclip_dwconv_slots(args, backward) */
  const void* dy;                 /* bf16 [n,ho,wo,c]                      (backward) */
  void* dx;                       /* bf16 [n,h,w,c] */
  float* dweight; int accumulate; /* fp32 [c,1,k,k] */
  float* dw_partials;             /* fp32 [stat_slots][k*k][c] workspace */
  float* bn_partials;             /* fp32 [stat_slots][2][c] (used iff in_scale != NULL) */
  const float* in_mean;           /* fp32 [c] batch statistics of the input BatchNorm */
  const float* in_invstd;
} mclip_dwconv_args;
int mclip_dwconv_slots(const mclip_dwconv_args* args, int backward);
int mclip_dwconv_forward(const mclip_dwconv_args* args, void* stream);
int mclip_dwconv_backward(const mclip_dwconv_args* args, void* stream);

/* mclip_stem_im2col: EfficientNet._conv_stem (efficientnet_custom.py:174-176,273), dense 3x3 stride-2 on 3 channels with
 * static pads, as im2col + tcgen05 GEMM.  Gathers each output pixel's 27 taps (t = ci*9 + ky*3 + kx, zero padded to 32)
 * from the fp32 image with arbitrary element strides (the trainer passes NCHW-shaped NHWC memory, trainer_ddp.py:288-291)
 * into bf16 rows; forward = mclip_gemm_tn(patches, W[c,32]) (+ BN statistics), weight gradient = mclip_gemm_wgrad(dY, patches). */
typedef struct mclip_stem_args {
  int n, h, w, ho, wo;
  int pad_left, pad_right, pad_top, pad_bottom;
  const void* in; long long stride_n, stride_c, stride_h, stride_w;   /* element strides; stride_c ignored when in_channels == 1 */
  void* out;                      /* bf16 [n*ho*wo, 32] */
  /* input edge (ABI 4): the data pipeline's grey-level image is a single channel replicated three times
   * (datasets/imagetext.py:121 `convert('RGB')`).  in_channels == 1 reads ONE channel and writes it to the three tap groups,
   * bit-identical to the 3-identical-channel input.  in_dtype: 0 fp32, 1 fp16, 2 bf16, 3 uint8.  With uint8 and norm_lut != NULL
   * the per-image normalisation of imagetext.py:129-134 is applied on load through a per-image 256-entry table (bf16
   * [n][256], from mclip_image_norm_lut_u8) that holds, for every grey level u, the reference's own fp32 result
   * ((u - min) / range - mean) / std rounded to the bf16 the patch stores anyway. */
  int in_channels, in_dtype;
  const void* norm_lut;
} mclip_stem_args;
int mclip_stem_im2col(const mclip_stem_args* args, void* stream);
/* uint8 [n, hw] image batch -> per-image {min, max - min} as fp32 [n][2] (imagetext.py:130-131) and the normalisation table
 * lut[n][u] = bf16(((u - min) / (max - min) - mean) / std) in the reference's fp32 operation order (IEEE division). */
int mclip_image_norm_lut_u8(const void* in, int n, long long hw, float mean, float std, float* minmax, void* lut, void* stream);

/* ---- folded BatchNorm backward of the expand convolution (ABI 4) ---------------------------------------------------
 * Autograd of MBConvBlock's `_bn0(_expand_conv(x))` (efficientnet_custom.py:105-106) normally applies
 *   dY0 = a (dV0 - c1 - yhat c2),  a = gamma*invstd, c1 = mean(dV0), c2 = mean(dV0*yhat)
 * in a pass over the 6x-wide tensor and then runs dX = dY0 We and dWe = dY0^T X.  Both consumers are LINEAR in dY0 and
 * Y0 = X We^T, so the correction folds into the operands (T = diag(t), t = -a c2 invstd, xbar = mean of X):
 *   dX  = [dV0 | X] [diag(a) We ; G] + bias,   G = We^T T We,   bias = -(a c1)^T We - xbar G
 *   dWe = diag(a) (dV0^T X - c1 (sum X)^T) + T We (X^T X - (sum X)(sum X)^T / count)
 * phase 0: We, BN vectors -> wcat[:, :k1pad] = bf16(diag(a) We)^T (zeros up to k1pad), twe = bf16(T We), bias = -(a c1)^T We
 * phase 1: G (fp32 [cin,cin] = twe^T We from mclip_gemm_wgrad), sum X -> wcat[:, k1pad:] = bf16(G), bias -= xbar bf16(G)
 * phase 2: X^T X (fp32 [cin,cin]), sum X -> gc = bf16(X^T X - (sum X)(sum X)^T / count)
 * phase 3: dwe (holding dV0^T X), q = bf16(We gc) -> dwe = a (dwe - c1 (sum X)^T) + t q                       */
typedef struct mclip_bn0_fold_args {
  int cexp, cin, k1pad; long long ldw; double count;
  const float* we;                                  /* fp32 [cexp, cin] */
  const float* scale; const float* invstd; const float* c1; const float* c2;     /* fp32 [cexp] */
  void* wcat; void* twe; float* bias;               /* bf16 [cin, ldw], bf16 [cexp, cin], fp32 [cin] */
  const float* g; const float* sumx;                /* fp32 [cin, cin], fp32 [cin] */
  void* gc;                                         /* bf16 [cin, cin] */
  float* dwe; const void* q;                        /* fp32 [cexp, cin] in/out, bf16 [cexp, cin] */
} mclip_bn0_fold_args;
int mclip_bn0_fold(const mclip_bn0_fold_args* args, int phase, void* stream);

/* ---- BatchNorm / swish / squeeze-excite / pooling passes ---------------------------------------------------------
 * Producer kernels (GEMM, depthwise, stem) emit per-channel (sum, sum sq) partials; mclip_bn_finalize turns them
 * into the affine a = gamma*invstd, b = beta - mean*a that CONSUMERS apply while loading, and performs the running
 * statistics update of nn.BatchNorm2d(momentum=0.01, eps=1e-3) (efficientnet_custom.py:53-54,64,74,88,177,205):
 * biased variance for normalisation, unbiased for running_var, num_batches_tracked += 1.  training == 0: the affine
 * comes from the running statistics (eval mode). */
typedef struct mclip_bn_args {
  int c, slots, training;
  long long count;                      /* elements per channel (N*H*W) */
  const float* partials;                /* fp32 [slots][2][c] */
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* num_batches_tracked;
  float momentum, eps;
  float* scale; float* shift; float* mean; float* invstd;     /* fp32 [c] out */
} mclip_bn_args;
int mclip_bn_finalize(const mclip_bn_args* args, void* stream);

/* One streaming pass over y[n,hw,c] (bf16): u = act(scale*y+shift) * rowscale[n] + residual ; optional bf16 output and
 * per-(n,chunk) channel sums (average-pool partials).  Covers bn->swish->avg_pool (efficientnet_custom.py:110-115,
 * 283,309) and _bn2 + drop_connect + skip (:123-131; rowscale = mask/keep_prob, efficient_net_custom_utils.py:145-154). */
typedef struct mclip_ew_args {
  int n, hw, c, act, chunks;
  const void* y; const float* scale; const float* shift;
  const float* rowscale; const void* residual;
  void* out; float* pool_partials;      /* pool_partials fp32 [n][chunks][c], chunks = mclip_ew_chunks(n,hw,c) */
} mclip_ew_args;
int mclip_ew_chunks(int n, int hw, int c);
/* Tuning switch of the streaming passes: bit m (0..2) = mclip_ew_backward mode m, bit 3 = mclip_ew_forward stage their
 * inputs through a per-thread cp.async ring in shared memory (more bytes in flight per SM) instead of registers.
 * Same results bit for bit; returns the previous mask.  Initial value: env MCLIP_EW_ASYNC, else the built-in default. */
int mclip_set_ew_async(int mask);
int mclip_ew_forward(const mclip_ew_args* args, void* stream);
int mclip_pool_finalize(const float* partials, int n, int chunks, int c, int hw, const float* mult, float* out, void* stream);

/* Squeeze-excite FC stack (efficientnet_custom.py:114-119) and its backward; gate folded into per-sample weights. */
typedef struct mclip_se_args {
  int n, hw, c, cse, chunks, accumulate;
  const float* pool_partials;           /* fp32 [n][chunks][c] (forward) */
  const float* w1; const float* b1;     /* _se_reduce: [cse,c], [cse] */
  const float* w2; const float* b2;     /* _se_expand: [c,cse], [c] */
  float* pooled; float* z1; float* gate;          /* fp32 [n,c], [n,cse], [n,c] (saved for backward) */
  const float* dgate_partials;          /* fp32 [n][chunks][stride] (backward): d gate sums; stride = dgate_chunk_stride or c */
  int dgate_chunk_stride;
  float* dz2; float* dz1; float* dpool;           /* fp32 [n,c], [n,cse], [n,c] */
  float* dw1; float* db1; float* dw2; float* db2;
} mclip_se_args;
int mclip_se_fc(const mclip_se_args* args, void* stream);
int mclip_se_fc_backward(const mclip_se_args* args, void* stream);
int mclip_se_scale_weights(const float* w, const float* gate, void* out_bf16, int n, int cout, int cexp, void* stream);

/* Backward streaming passes (autograd of BatchNorm2d + MemoryEfficientSwish (efficient_net_custom_utils.py:64-80) +
 * SE gating + drop-connect).  mode 0: BN-backward reduction partials [n*chunks][2][c] (sum dv, sum dv*yhat);
 * mode 1: dY = scale*(dv - c1 - yhat*c2) (bf16); mode 2: A2 = gate*swish(scale*y+shift) (bf16) + partials
 * [n][chunks][5][c] = (sum dU*u, sum dU*s', sum s', sum dU*s'*yhat, sum s'*yhat): the d-gate sums AND everything the
 * following BatchNorm backward needs (mclip_se_bn_combine), so no separate reduction pass runs over the tensor. */
typedef struct mclip_ew_bwd_args {
  int n, hw, c, act, mode, dv_given, chunks;
  const void* y; const float* scale; const float* shift;
  const void* du; const float* dvec; const float* gate; const float* dpool; const float* rowscale;
  const float* mean; const float* invstd; const float* c1; const float* c2;
  float* partials; void* out;
} mclip_ew_bwd_args;
int mclip_ew_backward(const mclip_ew_bwd_args* args, void* stream);
int mclip_se_bn_combine(const float* partials, int n, int chunks, int c, const float* gate, const float* dpool, float* out_n_2_c,
                        void* stream);
int mclip_bn_bwd_finalize(const float* partials, int slots, int c, long long count, int training, float* dgamma, float* dbeta,
                          int accumulate, float* c1, float* c2, void* stream);

/* ---- BERT text tower (non-GEMM pieces) ---------------------------------------------------------------------------
 * What HuggingfaceTextEncoder.forward (text_encoder.py:47-49) runs inside transformers' BertModel, post-LN BERT:
 * embeddings + LayerNorm(eps 1e-12) + dropout; softmax(QK^T/8 + padding mask) (dropout) V per head; LayerNorm.
 * All Linear layers go through mclip_gemm_tn (bias / erf-GELU / dropout mask / residual epilogues). */
typedef struct mclip_bert_embed_args {
  int batch, seq_len, hidden, vocab, max_positions;
  const void* input_ids; const void* token_type_ids;      /* int64 [batch, seq_len] (token_type_ids may be NULL) */
  const float* word; const float* pos; const float* type; /* fp32 embedding tables */
  const float* gamma; const float* beta; float eps;
  const void* dropmask; float drop_scale;                 /* uint8 [batch*seq_len, hidden] keep-mask or NULL; 1/(1-p) */
  void* out;                                              /* bf16 [batch*seq_len, hidden] */
} mclip_bert_embed_args;
int mclip_bert_embed_ln(const mclip_bert_embed_args* args, void* stream);
int mclip_layernorm(const void* x_bf16, const float* gamma, const float* beta, float eps, void* out_bf16, int rows, int hidden, void* stream);
/* qkv: bf16 [batch*seq_len, 3*heads*head_dim] (Q | K | V); attention_mask int64 [batch, seq_len] (1 = attend);
 * dropmask uint8 [batch, heads, seq_len, seq_len] or NULL; out bf16 [batch*seq_len, heads*head_dim];
 * lse fp32 [batch, heads, seq_len] or NULL: log-sum-exp of the scaled masked scores, saved for the backward pass.
 * seq_len <= 256 (head_dim 64): tcgen05 kernels (csrc/attention.cu: TMA-staged tiles, scores in TMEM, one thread per query row);
 * longer sequences: the SIMT kernels of csrc/bert.cu.  Replaces transformers' BertSelfAttention (SDPA, modeling_bert.py:168-207). */
int mclip_bert_attention(const void* qkv, const void* attention_mask, const void* dropmask, float drop_scale, void* out, float* lse,
                         int batch, int seq_len, int heads, int head_dim, void* stream);

/* ---- BERT text tower, backward (non-GEMM pieces) -----------------------------------------------------------------
 * The reference trains every BERT parameter (optimizer/__init__.py:23-31 over model.parameters(); trainer_ddp.py:298-300
 * calls backward through transformers' BertModel).  Linear data/weight gradients run on mclip_gemm_tn / mclip_gemm_wgrad /
 * mclip_colsum; these entry points are the autograd of LayerNorm, erf-GELU, softmax attention and the embeddings. */
/* LayerNorm backward over x_bf16 [rows, hidden] (the pre-LN input) and dy_bf16: dx (bf16), optionally also
 * dx_drop = dx o keep-mask * drop_scale (the gradient entering the dropped sub-layer output, BertSelfOutput/BertOutput),
 * dgamma/dbeta (+)= column sums via partials fp32 [slots][2][hidden], slots = mclip_layernorm_backward_slots(rows). */
int mclip_layernorm_backward_slots(int rows);
int mclip_layernorm_backward(const void* x_bf16, const void* dy_bf16, const float* gamma, float eps, const void* dropmask, float drop_scale,
                             void* dx_bf16, void* dx_drop_bf16, float* partials, int slots, float* dgamma, float* dbeta, int accumulate,
                             int rows, int hidden, void* stream);
/* erf-GELU on bf16 (BertIntermediate): y = gelu(x); dx = dy * gelu'(x).  n = element count, multiple of 8. */
int mclip_gelu_forward(const void* x_bf16, void* y_bf16, long long n, void* stream);
int mclip_gelu_backward(const void* dy_bf16, const void* x_bf16, void* dx_bf16, long long n, void* stream);
/* d qkv (bf16 [batch*seq_len, 3*heads*head_dim]) from d_out (bf16 [batch*seq_len, heads*head_dim]) and the saved lse;
 * probabilities are recomputed from Q, K and the lse; delta = sum_k P dP is the exact fp32 row sum (not the bf16-rounded <dO, O>);
 * delta_ws fp32 [batch, heads, seq_len] is the pre-pass buffer of the SIMT path (seq_len > 256; the tcgen05 path keeps delta in
 * registers); every output element is written exactly once (deterministic). */
int mclip_bert_attention_backward(const void* qkv, const void* d_out, const float* lse, const void* attention_mask, const void* dropmask,
                                  float drop_scale, float* delta_ws, void* dqkv, int batch, int seq_len, int heads, int head_dim, void* stream);
/* Embeddings backward (BertEmbeddings): LayerNorm backward of (dout o keep-mask*scale) with the pre-LN sum recomputed from
 * the tables, then word rows (first occurrence of an id sums all its tokens in order: deterministic, no atomics),
 * position rows < seq_len and token-type rows.  accumulate == 0 writes ONLY the touched rows: the caller provides zeroed
 * (or previously accumulated) tables. */
typedef struct mclip_bert_embed_bwd_args {
  int batch, seq_len, hidden, vocab, max_positions, n_types, slots, accumulate;
  const void* input_ids; const void* token_type_ids;      /* int64 [batch, seq_len] (token_type_ids may be NULL) */
  const float* word; const float* pos; const float* type; const float* gamma; float eps;
  const void* dropmask; float drop_scale;                 /* the forward's keep-mask [batch*seq_len, hidden] or NULL */
  const void* dout;                                       /* bf16 [batch*seq_len, hidden]: gradient of the embedding output */
  float* dv; float* partials;                             /* workspaces: fp32 [batch*seq_len, hidden], [slots][2][hidden] */
  float* dword; float* dpos; float* dtype; float* dgamma; float* dbeta;
} mclip_bert_embed_bwd_args;
int mclip_bert_embed_backward(const mclip_bert_embed_bwd_args* args, void* stream);

/* fp32 master weights -> bf16 GEMM operands (and their transposes for the data-gradient GEMMs), one launch per tower. */
typedef struct mclip_prep_entry { const void* src; void* dst; void* dst_t; int rows, cols, dst_ld, dst_t_ld; } mclip_prep_entry;   /* dst_ld / dst_t_ld: row strides of dst / dst_t (0: cols / rows) */
int mclip_weight_prep(const void* table_dev, int n_entries, void* stream);

/* CLIP head helpers: cast, x/||x|| (clip.py:90-91) forward/backward, Linear bias gradient. */
int mclip_cast_bf16(const float* in, void* out_bf16, long long n, void* stream);
int mclip_l2norm_forward(const void* x_bf16, float* e, float* norm, int rows, int d, void* stream);
int mclip_l2norm_backward(const float* e, const float* de, const float* norm, void* dx_bf16, int rows, int d, void* stream);
long long mclip_colsum_workspace_bytes(int rows, int cols);
int mclip_colsum(const void* x_bf16, float* out, int rows, int cols, long long ld, int accumulate, void* workspace, long long workspace_bytes,
                 void* stream);   /* out[c] (+)= sum_r x[r,c]; workspace: fp32 row-split partials (NULL: slow single-pass kernel) */

/* AdamW over a flat fp32 buffer (breastclip/optimizer/__init__.py:23-31: torch.optim.AdamW on all parameters). */
int mclip_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, long long step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCLIP_H_ */
