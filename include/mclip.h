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
  void* workspace; long long workspace_bytes;  /* >= mclip_loss_workspace_bytes() */
  float* out;                                  /* [2 + 2*n_pairs] fp32 */
  /* world > 1 only: symmetric (peer-mapped) gather buffers, see mclip_symm_* below */
  float* gathered[MCLIP_LOSS_MAX_TENSORS];     /* this rank's [W*B,D] gather buffer per tensor */
  float* const* peer_gathered;                 /* device table [W][n_tensors]: every rank's gather buffers */
  uint32_t* const* peer_flags;                 /* device table [W]: every rank's arrival counters (u32[W]) */
  const uint32_t* my_flags;                    /* this rank's arrival counters */
  long long epoch;                             /* 1,2,3,... identical on all ranks, +1 per call on this buffer set */
} mclip_loss_args;
long long mclip_loss_workspace_bytes(int world, int batch, int dim, int n_pairs);
int mclip_loss_grid(int world, int batch, int n_pairs);
int mclip_contrastive_loss(const mclip_loss_args* args, void* stream);

/* ---- tcgen05 GEMMs ---------------------------------------------------------------------------------------------
 * mclip_gemm_tn: D[b,m,n] = epi(sum_k A[b,m,k] B[b|0,n,k]), bf16 in/out, fp32 accumulate in TMEM.  This is the 1x1
 * "pointwise" convolution of MBConv in NHWC (efficientnet_custom.py:105 _expand_conv, :122 _project_conv, :283
 * _conv_head; m = pixel, k = Cin, n = Cout), its data gradient (B = transposed weight), and every nn.Linear of the
 * text tower / projection heads (text_encoder.py:48 -> transformers BertModel; projection.py:28).
 * epi: + bias[n]; act 1 = erf-GELU; + residual[b,m,n]; optional per-column (sum, sum of squares) partials of the
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

#ifdef __cplusplus
}
#endif
#endif /* MCLIP_H_ */
