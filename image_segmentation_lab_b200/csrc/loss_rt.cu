// Host dispatch of the register-tile kernels (loss_rt.cuh); the kernels are instantiated per dtype in
// loss_rt_{f32,bf16,f16}.cu.
#include "common.cuh"

namespace b200seg {

struct RtParams;
}
#include "loss_rt_params.cuh"

namespace b200seg {

int rt_run_f32(const RtParams& p, int kind, bool vec, cudaStream_t st);
int rt_run_bf16(const RtParams& p, int kind, bool vec, cudaStream_t st);
int rt_run_f16(const RtParams& p, int kind, bool vec, cudaStream_t st);

static int rt_by_dtype(int dtype, const RtParams& p, int kind, bool vec, cudaStream_t st) {
  switch (dtype) {
    case B200SEG_F32: return rt_run_f32(p, kind, vec, st);
    case B200SEG_BF16: return rt_run_bf16(p, kind, vec, st);
    case B200SEG_F16: return rt_run_f16(p, kind, vec, st);
  }
  set_error("unsupported logit dtype %d", dtype);
  return 1;
}

bool bulk_supported(const void* logits, const void* labels, const void* grad, int logit_dtype, int label_dtype, int C,
                    long long HW, bool has_pixel_weight);
int bulk_fused_dispatch(const b200seg_loss_fused_desc* d, cudaStream_t st);

int flat_fused_dispatch(const b200seg_loss_fused_desc* d, cudaStream_t st) {
  const b200seg_loss_desc* f = &d->fwd;
  B200SEG_REQUIRE(d->grad_logits != nullptr, "loss_fused: grad_logits is NULL");
  B200SEG_REQUIRE(!d->use_nvalid, "loss_fused: avg_non_ignore needs the two-pass path at label resolution");
  // 16-byte tileable problems take the bulk-copy pipeline (any C whose tile fits shared memory); the rest the
  // register-tile kernel (C <= 32)
  if (bulk_supported(f->logits, f->labels, d->grad_logits, f->logit_dtype, f->label_dtype, f->C, (long long)f->H * f->W,
                     f->pixel_weight != nullptr))
    return bulk_fused_dispatch(d, st);
  B200SEG_REQUIRE(f->C <= 32, "loss_fused: the label-resolution single pass holds at most 32 classes unless the problem is "
                  "16-byte tileable (got %d; query b200seg_loss_flat_single_ok first)", f->C);
  RtParams p = {};
  p.logits = f->logits; p.labels = f->labels; p.pw = f->pixel_weight; p.cw = f->ce_class_weight;
  p.ce_grad_out = d->grad_out; p.stats = reinterpret_cast<unsigned long long*>(f->stats); p.grad = d->grad_logits;
  p.ce_scale_host = d->grad_scale_host; p.label_dtype = f->label_dtype;
  p.N = f->N; p.C = f->C; p.HW = (long long)f->H * f->W;
  p.flags = f->flags | B200SEG_WANT_CE;
  p.ignore_index = f->ignore_index; p.acc_has_ignore = f->acc_has_ignore; p.acc_ignore = f->acc_ignore_index;
  p.dice_exponent = 2.f; p.lw = f->ce_loss_weight;
  const bool vec = (p.HW % 2 == 0) && aligned16(f->logits) && aligned16(f->labels) && aligned16(d->grad_logits) &&
                   (!p.pw || aligned16(p.pw));
  return rt_by_dtype(f->logit_dtype, p, 0, vec, st);
}

int tile_fwd_dispatch(const b200seg_loss_desc* d, cudaStream_t st) {
  B200SEG_REQUIRE(d->h == d->H && d->w == d->W, "dice forward needs logits at label resolution (resize first)");
  B200SEG_REQUIRE(d->C <= 32, "register-tile dice path holds at most 32 classes (got %d)", d->C);
  B200SEG_REQUIRE(d->dice_part != nullptr, "dice_part workspace is NULL");
  B200SEG_REQUIRE(d->dice_exponent > 0.f, "dice exponent must be > 0");
  RtParams p = {};
  p.logits = d->logits; p.labels = d->labels; p.pw = d->pixel_weight; p.cw = d->ce_class_weight;
  p.lse_out = (d->flags & B200SEG_WANT_LSE) ? d->lse : nullptr;
  p.loss_px = (d->flags & B200SEG_WANT_LOSS_PX) ? d->loss_px : nullptr;
  p.stats = reinterpret_cast<unsigned long long*>(d->stats);
  p.dice_part = d->dice_part;
  p.label_dtype = d->label_dtype;
  p.N = d->N; p.C = d->C; p.HW = (long long)d->H * d->W;
  p.flags = d->flags; p.ignore_index = d->ignore_index;
  p.acc_has_ignore = d->acc_has_ignore; p.acc_ignore = d->acc_ignore_index;
  p.dice_ignore = d->dice_ignore_index; p.dice_exponent = d->dice_exponent; p.lw = d->ce_loss_weight;
  const bool vec = (p.HW % 2 == 0) && aligned16(d->logits) && aligned16(d->labels) &&
                   (!p.pw || aligned16(p.pw)) && (!p.lse_out || aligned16(p.lse_out)) &&
                   (!p.loss_px || aligned16(p.loss_px));
  return rt_by_dtype(d->logit_dtype, p, 1, vec, st);
}

int tile_bwd_dispatch(const b200seg_loss_bwd_desc* d, cudaStream_t st) {
  B200SEG_REQUIRE(d->h == d->H && d->w == d->W, "dice backward needs logits at label resolution");
  B200SEG_REQUIRE(d->C <= 32, "register-tile dice path holds at most 32 classes (got %d)", d->C);
  B200SEG_REQUIRE(d->dice_coef != nullptr && d->lse != nullptr, "dice backward needs dice_coef and lse");
  RtParams p = {};
  p.logits = d->logits; p.labels = d->labels; p.pw = d->pixel_weight; p.cw = d->ce_class_weight;
  p.lse_in = d->lse; p.ce_grad_out = d->ce_grad_out; p.ce_grad_px = d->ce_grad_px;
  p.dice_coef = d->dice_coef; p.dice_grad_out = d->dice_grad_out; p.grad = d->grad_logits;
  p.ce_scale_host = d->ce_scale_host; p.ce_use_nvalid = d->ce_use_nvalid;
  p.stats = const_cast<unsigned long long*>(reinterpret_cast<const unsigned long long*>(d->stats));
  p.label_dtype = d->label_dtype;
  p.N = d->N; p.C = d->C; p.HW = (long long)d->H * d->W;
  p.flags = d->flags; p.ignore_index = d->ignore_index;
  p.dice_ignore = d->dice_ignore_index; p.dice_exponent = d->dice_exponent;
  const bool vec = (p.HW % 2 == 0) && aligned16(d->logits) && aligned16(d->labels) && aligned16(d->lse) &&
                   aligned16(d->grad_logits) && (!p.pw || aligned16(p.pw)) && (!p.ce_grad_px || aligned16(p.ce_grad_px));
  return rt_by_dtype(d->logit_dtype, p, 2, vec, st);
}

}  // namespace b200seg
