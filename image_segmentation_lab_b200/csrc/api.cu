// C-ABI entry points of libb200seg.so: argument validation, workspace zeroing and kernel dispatch.
// Validation mirrors the Python-side errors of the reference; everything that can be
// checked without touching device memory is checked here, before any launch.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace b200seg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return 3;
  }
  return 0;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (kernel, device): remembered per pair, so that the first
// launch on a second GPU of the same process opts in again, and safe from several host threads.
int ensure_dyn_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;
  int dev = 0;
  B200SEG_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  int& have = done[std::make_pair(func, dev)];
  if (have >= bytes) return 0;
  B200SEG_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  have = bytes;
  return 0;
}

int ce_fwd_dispatch(const b200seg_loss_desc* d, cudaStream_t st);
int ce_bwd_dispatch(const b200seg_loss_bwd_desc* d, cudaStream_t st);
int finalize_dispatch(const b200seg_finalize_desc* d, cudaStream_t st);
int tile_fwd_dispatch(const b200seg_loss_desc* d, cudaStream_t st);
int tile_bwd_dispatch(const b200seg_loss_bwd_desc* d, cudaStream_t st);
int up_fused_dispatch(const b200seg_loss_fused_desc* d, cudaStream_t st);
int dice_stream_fwd_dispatch(const b200seg_loss_desc* d, cudaStream_t st);
int dice_stream_bwd_dispatch(const b200seg_loss_bwd_desc* d, cudaStream_t st);
long long up_fused_workspace(int N, int C, int h, int w, int H, int W, int ac);
int up_combine_dispatch(const void* ws, void* grad, int dtype, int N, int C, int h, int w, float scale_host,
                        const float* grad_out, int use_nvalid, const uint64_t* stats, cudaStream_t st);
int scale_inplace_dispatch(void* x, int dtype, long long n, const float* g, cudaStream_t st);
int lovasz_fwd_dispatch(const b200seg_lovasz_desc* d, cudaStream_t st);
int lovasz_bwd_dispatch(const b200seg_lovasz_bwd_desc* d, cudaStream_t st);
long long lovasz_workspace_bytes(int N, int C, long long HW, int per_image, int pairs);
bool bulk_fwd_supported(const b200seg_loss_desc* d);
int bulk_fwd_dispatch(const b200seg_loss_desc* d, cudaStream_t st);
static bool bulk_fwd_enabled() {   // B200SEG_NO_BULK_FWD=1: the streaming forward (A/B measurements and the old-vs-new parity test)
  const char* e = getenv("B200SEG_NO_BULK_FWD");
  return !(e && e[0] == '1');
}
bool bulk_supported(const void* logits, const void* labels, const void* grad, int logit_dtype, int label_dtype, int C,
                    long long HW, bool has_pixel_weight);

static int check_shape(const char* who, int N, int C, int h, int w, int H, int W, int ldt, int ydt) {
  B200SEG_REQUIRE(N >= 0 && C >= 1 && h >= 1 && w >= 1 && H >= 1 && W >= 1, "%s: bad shape N=%d C=%d h=%d w=%d H=%d W=%d",
                  who, N, C, h, w, H, W);
  B200SEG_REQUIRE(ldt == B200SEG_F32 || ldt == B200SEG_BF16 || ldt == B200SEG_F16, "%s: unsupported logit dtype %d", who, ldt);
  B200SEG_REQUIRE(ydt >= B200SEG_L_U8 && ydt <= B200SEG_L_F64, "%s: unsupported label dtype %d", who, ydt);
  B200SEG_REQUIRE(N <= 65535, "%s: batch %d exceeds 65535", who, N);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// top-k counts (accuracy.py:41-60 for arbitrary topk / thresh): rank of the label's logit.
template <typename T>
__global__ void __launch_bounds__(256) topk_counts_kernel(const T* __restrict__ logits, const void* __restrict__ labels,
                                                          int label_dtype, int C, long long HW, int has_ignore,
                                                          long long ignore, int k0, int k1, int k2, int k3, int nk,
                                                          int has_thresh, float thresh, unsigned long long* counts) {
  __shared__ double sred[5 * 32];
  const int n = blockIdx.y;
  const long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int hit[4] = {0, 0, 0, 0};
  int valid = 0;
  if (px < HW) {
    const long long y = load_label(labels, label_dtype, (size_t)n * HW + px);
    const bool live = has_ignore ? (y != ignore) : true;
    if (live) {
      valid = 1;
      if (y >= 0 && y < (long long)C) {
        const T* base = logits + (size_t)n * C * HW + px;
        const float zy = to_float<T>(base[(size_t)y * HW]);
        int rank = 0;
        for (int c = 0; c < C; ++c) {
          const float z = to_float<T>(base[(size_t)c * HW]);
          rank += (z > zy) || (z == zy && c < (int)y);
        }
        const bool pass = has_thresh ? (zy > thresh) : true;
        const int ks[4] = {k0, k1, k2, k3};
#pragma unroll
        for (int j = 0; j < 4; ++j) hit[j] = (j < nk && pass && rank < ks[j]);
      }
    }
  }
  double r[5] = {(double)hit[0], (double)hit[1], (double)hit[2], (double)hit[3], (double)valid};
  block_sum<double, 5>(r, sred);
  if (threadIdx.x == 0) {
    for (int j = 0; j < nk; ++j)
      if (r[j] != 0.0) atomicAdd(counts + j, (unsigned long long)r[j]);
    if (r[4] != 0.0) atomicAdd(counts + nk, (unsigned long long)r[4]);
  }
}

}  // namespace b200seg

using namespace b200seg;

extern "C" const char* b200seg_last_error(void) { return g_err; }
extern "C" int32_t b200seg_abi_version(void) { return B200SEG_ABI_VERSION; }
extern "C" int64_t b200seg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int b200seg_loss_fwd(const b200seg_loss_desc* d, void* stream) {
  B200SEG_REQUIRE(d != nullptr, "loss_fwd: NULL descriptor");
  if (int e = check_shape("loss_fwd", d->N, d->C, d->h, d->w, d->H, d->W, d->logit_dtype, d->label_dtype)) return e;
  B200SEG_REQUIRE(d->stats != nullptr, "loss_fwd: NULL stats");
  B200SEG_REQUIRE((d->flags & (B200SEG_WANT_CE | B200SEG_WANT_DICE | B200SEG_WANT_ACC)) != 0, "loss_fwd: nothing requested");
  B200SEG_REQUIRE(!(d->flags & B200SEG_WANT_LSE) || d->lse, "loss_fwd: WANT_LSE without lse buffer");
  B200SEG_REQUIRE(!(d->flags & B200SEG_WANT_LOSS_PX) || d->loss_px, "loss_fwd: WANT_LOSS_PX without loss_px buffer");
  cudaStream_t st = (cudaStream_t)stream;
  B200SEG_CUDA(cudaMemsetAsync(d->stats, 0, B200SEG_STATS_WORDS * sizeof(uint64_t), st));
  if (d->flags & B200SEG_WANT_DICE) {
    B200SEG_REQUIRE(d->dice_part != nullptr, "loss_fwd: WANT_DICE without dice_part");
    B200SEG_CUDA(cudaMemsetAsync(d->dice_part, 0, (size_t)d->N * d->C * 3 * sizeof(double), st));
  }
  if (d->N == 0) return 0;
  B200SEG_REQUIRE(d->logits && d->labels, "loss_fwd: NULL logits/labels");
  // Dice: <= 32 classes fit one warp's register tile (single read); more classes take the streaming kernels
  if (d->flags & B200SEG_WANT_DICE) {
    if (d->dice_mode == B200SEG_MODE_TVERSKY) {   // masked sums, exponent 1: the streaming kernels for every C
      B200SEG_REQUIRE(d->dice_exponent == 1.f, "loss_fwd: Tversky mode needs dice_exponent == 1");
      B200SEG_REQUIRE(d->h == d->H && d->w == d->W, "loss_fwd: Tversky needs logits at label resolution (resize first)");
      return dice_stream_fwd_dispatch(d, st);
    }
    B200SEG_REQUIRE(d->dice_mode == B200SEG_MODE_DICE, "loss_fwd: unknown dice_mode %d", d->dice_mode);
    return d->C <= 32 ? tile_fwd_dispatch(d, st) : dice_stream_fwd_dispatch(d, st);
  }
  // 16-byte tileable problems whose class tile fits shared memory take the bulk-copy pipeline in its forward-only form
  if (bulk_fwd_enabled() && bulk_fwd_supported(d)) return bulk_fwd_dispatch(d, st);
  return ce_fwd_dispatch(d, st);
}

extern "C" int b200seg_loss_finalize(const b200seg_finalize_desc* d, void* stream) {
  B200SEG_REQUIRE(d != nullptr && d->stats && d->out_loss_ce, "loss_finalize: NULL argument");
  B200SEG_REQUIRE(!(d->ce_has_avg_factor && d->ce_reduction == B200SEG_RED_SUM),
                  "avg_factor can not be used with reduction=\"sum\"");  // models/losses/utils.py:78-79
  B200SEG_REQUIRE(!(d->dice_has_avg_factor && d->dice_reduction == B200SEG_RED_SUM),
                  "avg_factor can not be used with reduction=\"sum\"");
  return finalize_dispatch(d, (cudaStream_t)stream);
}

extern "C" int b200seg_loss_bwd(const b200seg_loss_bwd_desc* d, void* stream) {
  B200SEG_REQUIRE(d != nullptr, "loss_bwd: NULL descriptor");
  if (int e = check_shape("loss_bwd", d->N, d->C, d->h, d->w, d->H, d->W, d->logit_dtype, d->label_dtype)) return e;
  if (d->N == 0) return 0;
  B200SEG_REQUIRE(d->logits && d->labels && d->lse && d->grad_logits, "loss_bwd: NULL tensor");
  B200SEG_REQUIRE(!d->ce_use_nvalid || d->stats, "loss_bwd: ce_use_nvalid without stats");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->flags & B200SEG_WANT_DICE) {
    if (d->dice_mode == B200SEG_MODE_TVERSKY) return dice_stream_bwd_dispatch(d, st);
    return d->C <= 32 ? tile_bwd_dispatch(d, st) : dice_stream_bwd_dispatch(d, st);
  }
  B200SEG_REQUIRE(d->flags & B200SEG_WANT_CE, "loss_bwd: nothing requested");
  return ce_bwd_dispatch(d, st);
}

extern "C" int64_t b200seg_loss_fused_workspace_bytes(int32_t N, int32_t C, int32_t h, int32_t w, int32_t H, int32_t W,
                                                      int32_t align_corners) {
  return up_fused_workspace(N, C, h, w, H, W, align_corners);
}

extern "C" int32_t b200seg_loss_flat_single_ok(const void* logits, const void* labels, int32_t logit_dtype, int32_t label_dtype,
                                               int32_t C, int64_t HW, int32_t has_pixel_weight) {
  if (C <= 32) return 1;   // register-tile kernel
  return bulk_supported(logits, labels, logits, logit_dtype, label_dtype, C, HW, has_pixel_weight != 0) ? 1 : 0;
}

extern "C" int b200seg_loss_fused_fwdbwd(const b200seg_loss_fused_desc* d, void* stream) {
  B200SEG_REQUIRE(d != nullptr, "loss_fused: NULL descriptor");
  const b200seg_loss_desc* f = &d->fwd;
  if (int e = check_shape("loss_fused", f->N, f->C, f->h, f->w, f->H, f->W, f->logit_dtype, f->label_dtype)) return e;
  B200SEG_REQUIRE(!(f->flags & B200SEG_WANT_DICE), "loss_fused: dice is not supported by the single-pass entry");
  B200SEG_REQUIRE(f->stats != nullptr, "loss_fused: NULL stats");
  cudaStream_t st = (cudaStream_t)stream;
  B200SEG_CUDA(cudaMemsetAsync(f->stats, 0, B200SEG_STATS_WORDS * sizeof(uint64_t), st));
  if (f->N == 0) return 0;
  B200SEG_REQUIRE(f->logits && f->labels, "loss_fused: NULL logits/labels");
  return up_fused_dispatch(d, st);
}

extern "C" int b200seg_loss_fused_combine(const void* workspace, void* grad_logits, int32_t logit_dtype, int32_t N,
                                          int32_t C, int32_t h, int32_t w, float scale_host, const float* grad_out,
                                          int32_t use_nvalid, const uint64_t* stats, void* stream) {
  B200SEG_REQUIRE(workspace && grad_logits, "loss_fused_combine: NULL buffer");
  B200SEG_REQUIRE(N >= 0 && C >= 1 && h >= 1 && w >= 1, "loss_fused_combine: bad shape");
  B200SEG_REQUIRE(!use_nvalid || stats, "loss_fused_combine: use_nvalid without stats");
  if (N == 0) return 0;
  return up_combine_dispatch(workspace, grad_logits, logit_dtype, N, C, h, w, scale_host, grad_out, use_nvalid, stats,
                             (cudaStream_t)stream);
}

extern "C" int b200seg_scale_inplace(void* x, int32_t dtype, int64_t n, const float* g, void* stream) {
  B200SEG_REQUIRE(x && g && n >= 0, "scale_inplace: bad arguments");
  if (n == 0) return 0;
  return scale_inplace_dispatch(x, dtype, n, g, (cudaStream_t)stream);
}

extern "C" int64_t b200seg_lovasz_workspace_bytes(int32_t N, int32_t C, int64_t HW, int32_t per_image, int32_t pairs) {
  return lovasz_workspace_bytes(N, C, HW, per_image, pairs);
}

extern "C" int b200seg_lovasz_fwd(const b200seg_lovasz_desc* d, void* stream) {
  B200SEG_REQUIRE(d != nullptr, "lovasz_fwd: NULL descriptor");
  B200SEG_REQUIRE(d->N >= 0 && d->N <= 65535 && d->C >= 1 && d->HW >= 0, "lovasz_fwd: bad shape N=%d C=%d HW=%lld", d->N, d->C,
                  (long long)d->HW);
  B200SEG_REQUIRE(d->logit_dtype == B200SEG_F32 || d->logit_dtype == B200SEG_BF16 || d->logit_dtype == B200SEG_F16,
                  "lovasz_fwd: unsupported logit dtype %d", d->logit_dtype);
  B200SEG_REQUIRE(d->label_dtype >= B200SEG_L_U8 && d->label_dtype <= B200SEG_L_F64, "lovasz_fwd: unsupported label dtype %d",
                  d->label_dtype);
  B200SEG_REQUIRE(!d->binary || d->C == 1, "lovasz_fwd: the binary hinge takes single-channel logits (C=%d)", d->C);
  B200SEG_REQUIRE(d->binary || d->C <= 32766, "lovasz_fwd: at most 32766 classes");
  B200SEG_REQUIRE(d->reduction >= B200SEG_RED_NONE && d->reduction <= B200SEG_RED_SUM, "lovasz_fwd: bad reduction %d", d->reduction);
  B200SEG_REQUIRE(!(d->per_image && d->has_avg_factor && d->reduction == B200SEG_RED_SUM),
                  "avg_factor can not be used with reduction=\"sum\"");   // models/losses/utils.py:78-79
  B200SEG_REQUIRE(d->seg_stats && d->out, "lovasz_fwd: NULL seg_stats / out");
  B200SEG_REQUIRE((long long)d->N * d->HW < 2147483647LL, "lovasz_fwd: %lld pixels exceed 2^31-1", (long long)d->N * d->HW);
  if (d->N > 0 && d->HW > 0) {
    B200SEG_REQUIRE(d->logits && d->labels && d->lab16 && d->workspace, "lovasz_fwd: NULL tensor");
    B200SEG_REQUIRE(d->binary || d->lse, "lovasz_fwd: the multi-class loss needs the per-pixel log-sum-exp");
    B200SEG_REQUIRE((reinterpret_cast<uintptr_t>(d->workspace) & 255u) == 0, "lovasz_fwd: workspace must be 256-byte aligned");
    if (!d->binary && d->classes_host) {
      B200SEG_REQUIRE(d->n_classes >= 1, "lovasz_fwd: empty class list");
      for (int j = 0; j < d->n_classes; ++j)
        B200SEG_REQUIRE(d->classes_host[j] >= 0 && d->classes_host[j] < d->C, "lovasz_fwd: class %d outside [0,%d)",
                        d->classes_host[j], d->C);
    }
  }
  return lovasz_fwd_dispatch(d, (cudaStream_t)stream);
}

extern "C" int b200seg_lovasz_bwd(const b200seg_lovasz_bwd_desc* d, void* stream) {
  B200SEG_REQUIRE(d != nullptr, "lovasz_bwd: NULL descriptor");
  B200SEG_REQUIRE(d->N >= 0 && d->N <= 65535 && d->C >= 1 && d->HW >= 0, "lovasz_bwd: bad shape");
  B200SEG_REQUIRE(d->logit_dtype == B200SEG_F32 || d->logit_dtype == B200SEG_BF16 || d->logit_dtype == B200SEG_F16,
                  "lovasz_bwd: unsupported logit dtype %d", d->logit_dtype);
  if (d->N == 0 || d->HW == 0) return 0;
  B200SEG_REQUIRE(d->logits && d->lab16 && d->G && d->coef && d->grad_logits, "lovasz_bwd: NULL tensor");
  B200SEG_REQUIRE(d->binary || d->lse, "lovasz_bwd: NULL lse");
  return lovasz_bwd_dispatch(d, (cudaStream_t)stream);
}

extern "C" int b200seg_topk_counts(const void* logits, const void* labels, int32_t logit_dtype, int32_t label_dtype,
                                   int32_t N, int32_t C, int64_t HW, int32_t has_ignore, int64_t ignore_index,
                                   const int32_t* topk_host, int32_t n_topk, int32_t has_thresh, float thresh,
                                   int64_t* counts, void* stream) {
  B200SEG_REQUIRE(n_topk >= 1 && n_topk <= 4 && topk_host, "topk_counts: between 1 and 4 k values supported");
  B200SEG_REQUIRE(N >= 0 && N <= 65535 && C >= 1 && HW >= 1 && counts, "topk_counts: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  B200SEG_CUDA(cudaMemsetAsync(counts, 0, (size_t)(n_topk + 1) * sizeof(int64_t), st));
  if (N == 0) return 0;
  B200SEG_REQUIRE(logits && labels, "topk_counts: NULL tensor");
  int k[4] = {0, 0, 0, 0};
  for (int j = 0; j < n_topk; ++j) {
    B200SEG_REQUIRE(topk_host[j] >= 1 && topk_host[j] <= C, "maxk %d exceeds pred dimension %d", topk_host[j], C);
    k[j] = topk_host[j];
  }
  dim3 grid((unsigned)((HW + 255) / 256), N);
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counts);
  switch (logit_dtype) {
    case B200SEG_F32:
      topk_counts_kernel<float><<<grid, 256, 0, st>>>((const float*)logits, labels, label_dtype, C, HW, has_ignore,
                                                      ignore_index, k[0], k[1], k[2], k[3], n_topk, has_thresh, thresh, cnt);
      break;
    case B200SEG_BF16:
      topk_counts_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)logits, labels, label_dtype, C, HW,
                                                              has_ignore, ignore_index, k[0], k[1], k[2], k[3], n_topk,
                                                              has_thresh, thresh, cnt);
      break;
    case B200SEG_F16:
      topk_counts_kernel<__half><<<grid, 256, 0, st>>>((const __half*)logits, labels, label_dtype, C, HW, has_ignore,
                                                       ignore_index, k[0], k[1], k[2], k[3], n_topk, has_thresh, thresh, cnt);
      break;
    default:
      set_error("topk_counts: unsupported logit dtype %d", logit_dtype);
      return 1;
  }
  count_launch();
  return check_launch("topk_counts_kernel");
}
