/*
 * b200seg.h — C ABI of libb200seg.so: the B200 (sm_100a) logits -> loss -> metrics hot path.
 *
 * This is the drop-in boundary. The reference (HanHan-TR/Image_Segmentation_lab) has no
 * native code; on this path it calls ATen. Each entry point below replaces one chain of those
 * calls, cited as reference file:line:
 *
 *   b200seg_resize_bilinear_fwd/bwd  utils/ops.py:7-26 (F.interpolate, mode='bilinear'), hot call
 *                                    models/decode_heads/decode_head.py:266-269
 *   b200seg_loss_fwd / _bwd          decode_head.py:266-295 fused:
 *                                      resize              utils/ops.py:26
 *                                      cross_entropy       models/losses/cross_entropy_loss.py:23-74
 *                                      weight_reduce_loss  models/losses/utils.py:48-80
 *                                      DiceLoss.forward    models/losses/dice_loss.py:103-134 (+ :23-58)
 *                                      TverskyLoss.forward models/losses/tversky_loss.py:112-134 (+ :24-68)
 *                                      accuracy (top-1)    models/losses/accuracy.py:6-61
 *   b200seg_loss_fused_fwdbwd        the same chain plus its autograd backward in one pass
 *   b200seg_lovasz_fwd / _bwd        LovaszLoss.forward models/losses/lovasz_loss.py:272-298 (+ lovasz_grad :26-39,
 *                                    lovasz_softmax(_flat) :135-231, lovasz_hinge(_flat) :69-132) and its autograd backward
 *   b200seg_confusion_labels         SegEvaluator.intersect_and_union core/evaluation/metrics.py:210-270
 *   b200seg_confusion_logits         SegEvaluator.process argmax :101-107 + intersect_and_union
 *
 * Conventions
 *   - Plain pointers and sizes only. Every pointer is a DEVICE pointer unless marked "host".
 *   - The caller owns every buffer (inputs, outputs, workspaces). Nothing is allocated or freed.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises
 *     and every call is CUDA-graph capturable.
 *   - Return value: 0 on success, non-zero on error; b200seg_last_error() returns a thread-local
 *     message for the last failing call on the calling thread.
 *   - Tensors are dense, row-major: logits (N,C,h,w), labels / per-pixel maps (N,H,W).
 */
#ifndef B200SEG_H_
#define B200SEG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SEG_ABI_VERSION 14

/* logit element types */
enum { B200SEG_F32 = 0, B200SEG_BF16 = 1, B200SEG_F16 = 2 };
/* label element types (read directly; no .long() pass — reference cross_entropy_loss.py:283) */
enum { B200SEG_L_U8 = 0, B200SEG_L_I16 = 1, B200SEG_L_I32 = 2, B200SEG_L_I64 = 3, B200SEG_L_F32 = 4, B200SEG_L_F64 = 5 };

/* what b200seg_loss_fwd computes (bit flags) */
enum {
  B200SEG_WANT_CE      = 1,   /* softmax cross-entropy sum                                   */
  B200SEG_WANT_DICE    = 2,   /* per-(n,c) dice partial sums                                 */
  B200SEG_WANT_ACC     = 4,   /* top-1 correct / valid counts                                */
  B200SEG_WANT_LOSS_PX = 8,   /* write per-pixel CE loss (reduction='none')                  */
  B200SEG_WANT_LSE     = 16   /* write per-pixel log-sum-exp (saved for the backward)        */
};

/* Layout of the 64-bit statistics block filled by b200seg_loss_fwd (B200SEG_STATS_WORDS words). */
enum {
  B200SEG_ST_CE_SUM    = 0,   /* double : sum_px  pixel_weight * class_weight[y] * nll       */
  B200SEG_ST_N_VALID   = 1,   /* int64  : #pixels with label != ignore_index                 */
  B200SEG_ST_N_CORRECT = 2,   /* int64  : #valid pixels whose arg-max class == label         */
  B200SEG_ST_N_BAD     = 3,   /* int64  : #pixels whose label is neither ignore nor in [0,C) */
  B200SEG_ST_N_ACC     = 4,   /* int64  : accuracy denominator (honours acc_has_ignore)      */
  B200SEG_STATS_WORDS  = 8
};

/* layout of b200seg_finalize_desc.log_vec */
enum { B200SEG_LOG_CE_SUM = 0, B200SEG_LOG_N_VALID = 1, B200SEG_LOG_N_CORRECT = 2, B200SEG_LOG_N_ACC = 3,
       B200SEG_LOG_N_BAD = 4, B200SEG_LOG_N_PIXELS = 5, B200SEG_LOG_DICE_SUM = 6, B200SEG_LOG_N_IMAGES = 7,
       B200SEG_LOG_WORDS = 8 };

/* which overlap loss the B200SEG_WANT_DICE sums serve (dice_mode)
 *   DICE   : [sum p*t*v, sum p^e (unmasked), sum t (unmasked, labels clamped)]   models/losses/dice_loss.py:48-58
 *   TVERSKY: [TP = sum p*t*v, sum p*v, sum t*v] (all masked by v = label != ignore_index; exponent must be 1):
 *            FP = sum p*v - TP, FN = sum t*v - TP                                models/losses/tversky_loss.py:52-68
 *            the forward stores lse = +inf for ignored pixels, which masks them in every later pass            */
enum { B200SEG_MODE_DICE = 0, B200SEG_MODE_TVERSKY = 1 };

/* reductions (models/losses/utils.py:28-80) */
enum { B200SEG_RED_NONE = 0, B200SEG_RED_MEAN = 1, B200SEG_RED_SUM = 2 };

typedef struct b200seg_loss_desc {
  /* ---- tensors ---- */
  const void*  logits;          /* (N,C,h,w)  logit_dtype                                      */
  const void*  labels;          /* (N,H,W)    label_dtype                                      */
  const float* pixel_weight;    /* (N,H,W) f32 or NULL   cross_entropy(weight=)                */
  const float* ce_class_weight; /* (C) f32 or NULL       cross_entropy(class_weight=)          */
  int32_t logit_dtype, label_dtype;
  int32_t N, C, h, w, H, W;     /* (h,w) != (H,W) => bilinear resize fused in (utils/ops.py:26) */
  int32_t align_corners;
  int32_t flags;                /* B200SEG_WANT_*                                              */
  int64_t ignore_index;         /* CE ignore_index                                             */
  int32_t acc_has_ignore;       /* accuracy(ignore_index=None) -> 0                            */
  int32_t dice_mode;            /* B200SEG_MODE_DICE / _TVERSKY                                */
  int64_t acc_ignore_index;
  /* ---- dice (models/losses/dice_loss.py) ---- */
  int64_t dice_ignore_index;
  float   dice_exponent;
  float   reserved1;
  /* ---- outputs ---- */
  float*   lse;                 /* (N,H,W) f32 or NULL                                         */
  float*   loss_px;             /* (N,H,W) f32 or NULL : loss_weight*pixel_weight*cw[y]*nll    */
  float    ce_loss_weight;      /* applied to loss_px only                                     */
  float    reserved2;
  uint64_t* stats;              /* B200SEG_STATS_WORDS x 8 bytes, zeroed by the call           */
  double*  dice_part;           /* (N,C,3) doubles [sum p*t*v, sum p^e, sum t] zeroed by call  */
} b200seg_loss_desc;

/* Fused forward: resize + log-softmax + NLL (+ dice partial sums) (+ top-1 accuracy counts). */
int b200seg_loss_fwd(const b200seg_loss_desc* d, void* stream);

typedef struct b200seg_finalize_desc {
  const uint64_t* stats;
  const double*   dice_part;        /* or NULL */
  const float*    dice_class_weight;/* (C) or NULL */
  int32_t N, C;
  int64_t n_pixels;                 /* N*H*W                                                   */
  int32_t ce_reduction;             /* B200SEG_RED_MEAN / _SUM (NONE is handled by loss_px)    */
  int32_t ce_avg_non_ignore;        /* cross_entropy_loss.py:67-68                             */
  int32_t ce_has_avg_factor;        /* utils.py:72-76                                          */
  int32_t dice_has_avg_factor;
  double  ce_avg_factor;
  double  dice_avg_factor;
  float   ce_loss_weight;
  float   dice_loss_weight;
  float   dice_smooth;
  int32_t dice_reduction;
  int64_t dice_ignore_index;
  /* one float each, separately allocated by the caller (so that a framework can hand them out as independent
   * tensors: `loss[name] += ...` in decode_head.py:290 needs loss scalars that are not views of one buffer)   */
  float*  out_loss_ce;              /* loss_weight * reduced CE (0 written when ce_reduction is NONE: use loss_px) */
  float*  out_loss_dice;            /* or NULL                                                 */
  float*  out_acc;                  /* top-1 accuracy in percent (accuracy.py:55-60), or NULL  */
  float*  dice_coef;                /* (N,C,2) f32 [alpha,beta] for the backward, or NULL      */
  double* log_vec;                  /* B200SEG_LOG_WORDS doubles or NULL: the additive quantities of
                                     * this call as float64 (exact below 2^53), i.e. the payload of the
                                     * single per-step all-reduce across data-parallel ranks           */
  int32_t dice_mode;                /* B200SEG_MODE_*                                          */
  float   tversky_alpha;            /* weight of the false positives (tversky_loss.py:66)      */
  float   tversky_beta;             /* weight of the false negatives                           */
  int32_t reserved0;
} b200seg_finalize_desc;

/* One tiny launch: statistics -> loss_ce, loss_dice, acc_seg scalars (+ dice backward table). */
int b200seg_loss_finalize(const b200seg_finalize_desc* d, void* stream);

typedef struct b200seg_loss_bwd_desc {
  const void*  logits;
  const void*  labels;
  const float* pixel_weight;
  const float* ce_class_weight;
  const float* lse;                 /* (N,H,W) from the forward                                */
  int32_t logit_dtype, label_dtype;
  int32_t N, C, h, w, H, W;
  int32_t align_corners;
  int32_t flags;                    /* B200SEG_WANT_CE | B200SEG_WANT_DICE                     */
  int64_t ignore_index;
  int64_t dice_ignore_index;
  float   dice_exponent;
  /* CE coefficient: grad_z = G * pixel_weight * cw[y] * (p - onehot), with
   *   G = ce_scale_host * (*ce_grad_out or 1) / (ce_use_nvalid ? stats[N_VALID] + eps : 1)
   * per-pixel upstream gradient (reduction='none') multiplies in through ce_grad_px.            */
  float   ce_scale_host;
  const float*    ce_grad_out;      /* scalar f32 (device) or NULL                             */
  const float*    ce_grad_px;       /* (N,H,W) f32 or NULL                                     */
  const uint64_t* stats;            /* for n_valid when ce_use_nvalid                          */
  int32_t ce_use_nvalid;
  int32_t dice_mode;                /* B200SEG_MODE_*                                          */
  const float* dice_coef;           /* (N,C,2) from finalize                                   */
  const float* dice_grad_out;       /* scalar f32 (device) or NULL                             */
  void*   grad_logits;              /* (N,C,h,w) logit_dtype, fully overwritten                */
  float*  reserved_scratch;         /* unused (was the atomicAdd accumulator of the removed scatter backward) */
  float*  scratch_px;               /* (N,H,W) f32 scratch: required for dice with C > 32      */
} b200seg_loss_bwd_desc;

/* Backward at label resolution ((h,w) == (H,W)): d(loss)/d(logits) in one pass (softmax Jacobian, dice). The resize-fused
 * backward is b200seg_loss_fused_fwdbwd + _combine; there is no atomicAdd scatter path. */
int b200seg_loss_bwd(const b200seg_loss_bwd_desc* d, void* stream);

/* 1 if b200seg_loss_fused_fwdbwd can run a label-resolution problem (h == H, w == W) in a single pass: C <= 32 (register
 * tile), or up to 34 (fp32) / 69 (16-bit) classes on the bulk-copy pipeline when the tensors are 16-byte tileable (aligned
 * pointers, H*W*elem % 16 == 0, no per-pixel weight). grad_logits is assumed to be aligned like logits. */
int32_t b200seg_loss_flat_single_ok(const void* logits, const void* labels, int32_t logit_dtype, int32_t label_dtype,
                                    int32_t C, int64_t HW, int32_t has_pixel_weight);

/* Bytes of scratch needed by b200seg_loss_fused_fwdbwd for the given problem (0 if none). */
int64_t b200seg_loss_fused_workspace_bytes(int32_t N, int32_t C, int32_t h, int32_t w, int32_t H, int32_t W,
                                           int32_t align_corners);

typedef struct b200seg_loss_fused_desc {
  b200seg_loss_desc fwd;            /* CE (+ACC) only; dice is not supported by this entry     */
  /* grad_logits = scale * d(sum_px pw*cw*nll)/dz with
   *   scale = grad_scale_host * (*grad_out or 1) / (use_nvalid ? n_valid + eps : 1)
   * The upstream gradient is usually unknown while the forward runs: pass grad_out = NULL and
   * apply it later with b200seg_loss_fused_combine (resize-fused case, defer_combine = 1) or
   * b200seg_scale_inplace (label-resolution case).                                            */
  float   grad_scale_host;
  int32_t use_nvalid;               /* resize-fused case only                                  */
  const float* grad_out;            /* scalar f32 (device) or NULL                             */
  void*   grad_logits;              /* (N,C,h,w) logit_dtype; NULL = forward only (resize-fused) */
  void*   workspace;                /* b200seg_loss_fused_workspace_bytes()                    */
  int32_t defer_combine;            /* resize-fused: leave the corner sums in `workspace`      */
  int32_t reserved;
} b200seg_loss_fused_desc;

/* Forward and backward of resize + CE in ONE pass over the logits (the gradient of a sum/mean
 * loss does not depend on other pixels, so it is produced while the logits are on chip).      */
int b200seg_loss_fused_fwdbwd(const b200seg_loss_fused_desc* d, void* stream);

/* Second half of the resize-fused single pass: adds the 4 corner sums around every low-res logit
 * (fixed order, deterministic) and applies scale_host * (*grad_out or 1) / (n_valid + eps)?.  */
int b200seg_loss_fused_combine(const void* workspace, void* grad_logits, int32_t logit_dtype, int32_t N, int32_t C,
                               int32_t h, int32_t w, float scale_host, const float* grad_out, int32_t use_nvalid,
                               const uint64_t* stats, void* stream);

/* x *= *g in place (n elements of dtype); returns immediately on the device when *g == 1. */
int b200seg_scale_inplace(void* x, int32_t dtype, int64_t n, const float* g, void* stream);

/* Sigmoid cross-entropy, one-hot expansion fused in: binary_cross_entropy + _expand_onehot_labels,
 * models/losses/cross_entropy_loss.py:77-164 (CrossEntropyLoss(use_sigmoid=True), the shipped default config).  */
typedef struct b200seg_bce_desc {
  const void*  logits;          /* (N,C,HW) logit_dtype                                           */
  const void*  labels;          /* (N,HW)   label_dtype                                           */
  const float* pixel_weight;    /* (N,HW) f32 or NULL                                             */
  const float* pos_weight;      /* (C) f32 or NULL  (class_weight -> pos_weight, :160-161)        */
  int32_t logit_dtype, label_dtype;
  int32_t N, C;
  int64_t HW;
  int64_t ignore_index;
  int32_t single_channel;       /* prediction was (N,1,H,W): target = label (0/1), :126-134       */
  int32_t use_nvalid;           /* backward: divide by (n_valid_pixels*C + eps)  (avg_non_ignore) */
  float   loss_weight;          /* forward: multiplies loss_elem                                  */
  float   grad_scale_host;      /* backward: G = grad_scale_host * (*grad_out or 1) [/ n_valid]   */
  float*  loss_elem;            /* forward: (N,C,HW) f32 per-element loss or NULL                 */
  const float* grad_out;        /* backward: scalar f32 (device) or NULL                          */
  const float* grad_elem;       /* backward: (N,C,HW) f32 upstream gradient (reduction='none')    */
  void*   grad_logits;          /* backward: (N,C,HW) logit_dtype. FORWARD with grad_logits != NULL = single pass: the
                                 * gradient grad_scale_host * d(sum of losses)/d logits is written too (upstream gradient
                                 * taken as 1: rescale with b200seg_scale_inplace); needs use_nvalid == 0, loss_elem NULL */
  uint64_t* stats;              /* 8 words: [0] double: sum of weighted losses; [1] int64: valid pixels; [2] CTAs done;
                                 * [3] top-1 hits; [4] pixels counted by the accuracy                  */
  float*  out;                  /* forward: f32 scalar = out_scale_host * sum [/ (n_valid*C + eps) when use_nvalid],
                                 * written on the device by the last CTA; or NULL                  */
  float   out_scale_host;
  int32_t acc_has_ignore;       /* accuracy(..., ignore_index=None) -> 0                           */
  float*  acc_out;              /* forward: (1,) f32 top-1 accuracy of the same launch (accuracy.py:41-60: arg-max over the
                                 * class logits == label, over pixels with label != acc_ignore_index), or NULL */
  int64_t acc_ignore_index;
} b200seg_bce_desc;
int b200seg_bce_fwd(const b200seg_bce_desc* d, void* stream);   /* zeroes stats, then accumulates     */
int b200seg_bce_bwd(const b200seg_bce_desc* d, void* stream);   /* reads stats[1] when use_nvalid     */

/* Lovasz-Softmax ('multi_class') and Lovasz hinge ('binary') losses, models/losses/lovasz_loss.py:26-298.
 * One SEGMENT per (image group, class): the whole batch, or one image when per_image. Per segment: sort keys from the
 * logits row and the per-pixel log-sum-exp, a hand-written segmented radix sort (descending errors, stable; all
 * segments of a batch of classes per launch), a scan of the foreground bits in sorted order, the Jaccard increments in closed form, loss_c = sum e_i g_i and dloss_c/dp_c scattered into G.      */
typedef struct b200seg_lovasz_desc {
  const void*  logits;          /* multi-class: (N,C,HW) RAW logits (the soft-max of :281-282 is fused in);
                                 * binary: (N,HW) logits (C must be 1)                                              */
  const void*  labels;          /* (N,HW) label_dtype                                                               */
  const float* lse;             /* (N,HW) f32 per-pixel log-sum-exp of the logits (b200seg_loss_fwd with
                                 * B200SEG_WANT_LSE); unused (may be NULL) for binary                                */
  const float* class_weight;    /* (C) f32 or NULL (:165-166); ignored for binary (a placeholder there, :103-104)   */
  int32_t logit_dtype, label_dtype;
  int32_t N, C;
  int64_t HW;
  int64_t ignore_index;
  int32_t has_ignore;           /* ignore_index=None -> 0 (:44-45, :61-62)                                          */
  int32_t binary;               /* loss_type == 'binary'                                                            */
  int32_t per_image;            /* :119-126, :217-226                                                               */
  int32_t only_present;         /* classes == 'present' (:153-154); 0 for 'all' or an explicit list                 */
  const int32_t* classes_host;  /* HOST array of class ids to average (classes=[...]) or NULL = every class         */
  int32_t n_classes;
  int32_t reduction;            /* B200SEG_RED_* — used only when per_image (weight_reduce_loss over the images)    */
  int32_t has_avg_factor;
  float   loss_weight;
  double  avg_factor;
  int16_t* lab16;               /* out (N,HW): compact class ids (-1 ignored), kept for the backward                */
  float*   G;                   /* out: multi-class (C,N,HW) f32 — CLASS-major —, binary (N,HW) f32: dloss_seg/dp (resp. /dz), unscaled;
                                 * NULL = forward only (the sort then moves keys only)                              */
  void*    workspace;           /* b200seg_lovasz_workspace_bytes(...) bytes, 256-byte aligned                      */
  int64_t  workspace_bytes;
  double*  seg_stats;           /* (n_groups, C or 1, 2) doubles [loss, #foreground + 1], zeroed by the call        */
  float*   out;                 /* per_image && reduction none: n_groups floats; else 1 float (loss_weight applied) */
  float*   coef;                /* (n_groups, C or 1) f32: d out / d loss_seg, for the backward; or NULL            */
} b200seg_lovasz_desc;
/* Bytes the forward wants for this shape (C = 1 for binary; pairs = 1 when G != NULL or binary): the (key, pixel index)
 * double buffers of every (class, image group) segment, the digit histograms and the look-back descriptors of the radix
 * sort. Up to 4 GiB it covers all classes in one batch of launches; beyond that the classes go through in several
 * batches. b200seg_lovasz_fwd adapts to whatever workspace_bytes it is given (at least one class must fit). */
int64_t b200seg_lovasz_workspace_bytes(int32_t N, int32_t C, int64_t HW, int32_t per_image, int32_t pairs);
int b200seg_lovasz_fwd(const b200seg_lovasz_desc* d, void* stream);

typedef struct b200seg_lovasz_bwd_desc {
  const void*    logits;
  const float*   lse;           /* multi-class only */
  const int16_t* lab16;
  const float*   G;
  const float*   coef;
  const float*   grad_out;      /* device f32: scalar, or one per image group when grad_per_group; NULL = 1         */
  void*          grad_logits;   /* same shape and dtype as logits, fully overwritten (0 on ignored pixels)          */
  int32_t logit_dtype, N, C;
  int32_t binary, per_image, grad_per_group;
  int64_t HW;
} b200seg_lovasz_bwd_desc;
int b200seg_lovasz_bwd(const b200seg_lovasz_bwd_desc* d, void* stream);

/* Bilinear resize, ATen semantics (torch/include/ATen/native/UpSample.h:271-312,442-476); the forward is bit-identical to
 * F.interpolate on CUDA (FMA contraction pinned). scale_h / scale_w: > 0 = the source-index scale to use instead of in/out —
 * F.interpolate(scale_factor=s) uses (float)(1.0 / s) (compute_scales_value); <= 0 = from the sizes. Ignored when
 * align_corners != 0, as ATen does. */
int b200seg_resize_bilinear_fwd(const void* in, void* out, int32_t dtype, int32_t NC, int32_t h, int32_t w,
                                int32_t H, int32_t W, int32_t align_corners, float scale_h, float scale_w, void* stream);
/* Deterministic transpose (gather form): grad_in (NC,h,w) <- grad_out (NC,H,W). */
int b200seg_resize_bilinear_bwd(const void* grad_out, void* grad_in, int32_t dtype, int32_t NC, int32_t h,
                                int32_t w, int32_t H, int32_t W, int32_t align_corners, float scale_h, float scale_w,
                                void* stream);
/* Nearest resize (F.interpolate default mode of utils/ops.py:7-26) and its deterministic gather backward. */
int b200seg_resize_nearest_fwd(const void* in, void* out, int32_t dtype, int32_t NC, int32_t h, int32_t w,
                               int32_t H, int32_t W, float scale_h, float scale_w, void* stream);
int b200seg_resize_nearest_bwd(const void* grad_out, void* grad_in, int32_t dtype, int32_t NC, int32_t h, int32_t w,
                               int32_t H, int32_t W, float scale_h, float scale_w, void* stream);

/* One image of an evaluation batch. */
typedef struct b200seg_image {
  const void* pred;                 /* label map (H,W) pred_dtype, or logits (C,h,w) logit dtype */
  const void* gt;                   /* (H,W) gt_dtype                                          */
  int64_t n_pixels;                 /* H*W                                                     */
  int32_t h, w;                     /* logits only: source size (== H,W unless resize fused)   */
  int32_t H, W;
} b200seg_image;

/* Area histograms from label maps. areas: (n_images,3,C) int64 [intersect, pred, label],
 * ACCUMULATED into (caller zeroes). `images` is a device array of n_images descriptors.
 * `chunk_prefix` is a device array of n_images+1 int64: prefix sum of ceil(n_pixels/chunk).
 * totals_only != 0: `areas` is (3,C), the sum over all images (what seg_metrics needs, metrics.py:163-166) — every CTA
 * flushes its counters once instead of once per image it touches.                             */
int b200seg_confusion_labels(const b200seg_image* images, const int64_t* chunk_prefix, int32_t n_images,
                             int64_t total_chunks, int32_t chunk_pixels, int32_t pred_dtype, int32_t gt_dtype,
                             int32_t C, int64_t ignore_index, int64_t* areas, int32_t totals_only, void* stream);

/* Same, with the arg-max over classes fused in (logits (C,H,W) per image; lowest index wins ties);
 * optionally writes the int64 label map to pred_out[i] ((H,W) each, may be NULL).              */
int b200seg_confusion_logits(const b200seg_image* images, const int64_t* chunk_prefix, int32_t n_images,
                             int64_t total_chunks, int32_t chunk_pixels, int32_t logit_dtype, int32_t gt_dtype,
                             int32_t C, int64_t ignore_index, int64_t* areas, int64_t* const* pred_out,
                             int32_t totals_only, void* stream);
/* Same, with the bilinear resize of low-resolution logits (C,h,w) to the ground-truth size (H,W) fused into the
 * arg-max (decode_head.py:297-320 + metrics.py:101-107): image.h/w = logit size, image.H/W = ground-truth size,
 * image.n_pixels = H*W. The rescaled logits are never materialised.                                            */
int b200seg_confusion_logits_resized(const b200seg_image* images, const int64_t* chunk_prefix, int32_t n_images,
                                     int64_t total_chunks, int32_t chunk_pixels, int32_t logit_dtype, int32_t gt_dtype,
                                     int32_t C, int64_t ignore_index, int32_t align_corners, int64_t* areas,
                                     int64_t* const* pred_out, int32_t totals_only, void* stream);
int32_t b200seg_confusion_chunk_pixels(void);

/* top-k accuracy counts (models/losses/accuracy.py:6-61) for arbitrary k and thresh:
 * counts[j] = #valid pixels whose label ranks < topk[j]; counts[n_topk] = #valid pixels.      */
int b200seg_topk_counts(const void* logits, const void* labels, int32_t logit_dtype, int32_t label_dtype,
                        int32_t N, int32_t C, int64_t HW, int32_t has_ignore, int64_t ignore_index,
                        const int32_t* topk_host, int32_t n_topk, int32_t has_thresh, float thresh,
                        int64_t* counts, void* stream);

const char* b200seg_last_error(void);
int32_t b200seg_abi_version(void);
/* number of kernels this library has launched from the calling process (for bench accounting) */
int64_t b200seg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200SEG_H_ */
