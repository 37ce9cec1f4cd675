// Parameter block of the register-tile kernels (loss_rt.cuh).
#pragma once
namespace b200seg {
struct RtParams {
  const void* logits;
  const void* labels;
  const float* pw;
  const float* cw;
  float* lse_out;
  float* loss_px;
  unsigned long long* stats;
  double* dice_part;
  // backward / single pass
  const float* lse_in;
  const float* ce_grad_out;
  const float* ce_grad_px;
  const float* dice_coef;
  const float* dice_grad_out;
  void* grad;
  float ce_scale_host;
  int ce_use_nvalid;
  //
  int label_dtype;
  int N, C;
  long long HW;
  int flags;
  long long ignore_index;
  int acc_has_ignore;
  long long acc_ignore;
  long long dice_ignore;
  float dice_exponent;
  float lw;
  int tiles;
};
}  // namespace b200seg
