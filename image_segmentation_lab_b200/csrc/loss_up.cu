// Host side of the resize-fused single pass (resize + soft-max cross-entropy + accuracy, forward AND backward, on
// LOW-RESOLUTION logits) and its two small companion kernels, sm_100a.
//
// For a 'mean' / 'sum' reduction the gradient of the loss w.r.t. a pixel's logits depends on that pixel only (up to
// one global scale), so it is produced while the interpolated logits are still in registers: the (N,C,H,W) tensor is
// never materialised in either direction, instead of the reference's ~10 passes over it (utils/ops.py:26,
// cross_entropy_loss.py:56-61, accuracy.py:41 and their autograd backwards).
//
// Scope: logits at lower resolution than the labels (H >= h, W >= w), any ratio, both align_corners settings, C <= 512.
//
//   up_gen_kernel      (loss_upgen.cuh)   the single pass (thread per cell on packed fp32 math); writes each (band, run)
//                                         cell's 4 corner gradient sums to PB[n][c][band][run] (float4), deterministic,
//                                         no atomics. C > 32: forward launch + up_gen_bwd_tile_kernel per 32 classes
//   up_cell_kernel     (loss_upcell.cuh)  its round-1 predecessor (quad per cell, power-of-two scales): B200SEG_UPCELL=old
//   up_combine_kernel  (here)             adds the 4 cells around every low-resolution logit and applies the global scale
//                                         (upstream gradient, loss_weight, 1/denominator) -> grad_logits
//   scale_inplace_kernel (here)           late scaling of an already produced gradient (flat single-pass plan)
#include <cstdlib>

#include "common.cuh"

namespace b200seg {

// grad[n][c][y][x] = G * (cell(y,x).c11 + cell(y,x+1).c10 + cell(y+1,x).c01 + cell(y+1,x+1).c00)
// Every cell is read from L2 exactly once: a warp owns 31 consecutive columns x 8 rows of one (n,c) plane; lane l loads
// the cells of column x0 + l row by row (coalesced float4), its right neighbour's cell arrives by shuffle, and the row
// above stays in registers. (The first version read the 4 cells of every logit separately: 4x the L2 traffic for the
// 20 MB corner buffer, 10.8 us at config 2.)
constexpr int kCombineRows = 8;
template <typename T>
__global__ void __launch_bounds__(256) up_combine_kernel(const float* __restrict__ pb, T* __restrict__ grad, int NC, int h,
                                                         int w, float scale_host, const float* grad_out, int use_nvalid,
                                                         const unsigned long long* stats) {
  pdl_wait();
  float G = scale_host;
  if (grad_out) G *= __ldg(grad_out);
  if (use_nvalid) {
    const double nv = (double)(long long)stats[B200SEG_ST_N_VALID];
    G = (float)((double)G / (nv + 1.1920928955078125e-07));
  }
  const int lane = threadIdx.x & 31;
  const int xt = (w + 30) / 31, yt = (h + kCombineRows - 1) / kCombineRows;
  const long long tasks = (long long)NC * yt * xt;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long t = warp0; t < tasks; t += nwarps) {
    const int tx = (int)(t % xt);
    const long long r = t / xt;
    const int ty = (int)(r % yt);
    const long long nc = r / yt;
    const int x = tx * 31 + lane;                 // cell column of this lane; logit column too for lanes 0..30
    const int y0 = ty * kCombineRows;
    const int ny = min(kCombineRows, h - y0);
    const float4* cells = reinterpret_cast<const float4*>(pb) + nc * (long long)(h + 1) * (w + 1);
    const bool cin = x <= w;                      // cell columns run 0..w
    // all rows of the strip are requested before the first is used (9 independent 16-byte loads in flight per lane)
    float4 c[kCombineRows + 1];
#pragma unroll
    for (int i = 0; i <= kCombineRows; ++i)
      c[i] = (cin && i <= ny) ? __ldg(cells + (long long)(y0 + i) * (w + 1) + x) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kCombineRows; ++i) {
      // cell (y,x).c11 = up.w | cell (y,x+1).c10 = right neighbour's up.z | cell (y+1,x).c01 = dn.y | cell (y+1,x+1).c00 = right dn.x
      const float up_r_z = __shfl_down_sync(0xffffffffu, c[i].z, 1);
      const float dn_r_x = __shfl_down_sync(0xffffffffu, c[i + 1].x, 1);
      if (i < ny && lane < 31 && x < w)
        grad[(nc * h + y0 + i) * (long long)w + x] = from_float<T>(G * ((c[i].w + up_r_z) + (c[i + 1].y + dn_r_x)));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scale_inplace_kernel(T* __restrict__ x, long long n, const float* __restrict__ g) {
  const float s = __ldg(g);
  if (s == 1.f) return;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = from_float<T>(to_float<T>(x[i]) * s);
}

// the quad-per-cell kernel of round 1 (loss_upcell.cuh): integer power-of-two scale in [4,32], align_corners=False
static bool up_pow2_ok(int C, int h, int w, int H, int W, int ac, int* S_out) {
  if (ac) return false;
  if (h <= 0 || w <= 0 || H % h || W % w) return false;
  const int S = H / h;
  if (S != W / w || S < 4 || S > 32 || (S & (S - 1))) return false;
  if (C > 32) return false;
  *S_out = S;
  return true;
}
// the thread-per-cell kernels (loss_upgen.cuh): any up-sampling ratio, both align_corners settings; C <= 32 in one launch,
// up to 512 classes as one forward launch + one backward launch per tile of 32 classes
static bool up_fast_ok(int C, int h, int w, int H, int W, int ac) {
  (void)ac;
  if (h <= 0 || w <= 0 || C < 1 || C > 512) return false;
  if (H < h || W < w || (H == h && W == w)) return false;
  return true;
}
// Which kernel: the thread-per-cell kernel (loss_upgen.cuh) everywhere — 52.6 us at config 2 against 80.1 us for the
// quad-per-cell kernel of round 1 (loss_upcell.cuh), and the only one for other ratios / align_corners=True / C > 32.
// B200SEG_UPCELL=old runs the round-1 kernel where it applies (A/B measurements, tests).
static bool up_use_old(int C, int h, int w, int H, int W, int ac, int* S_out) {
  if (!up_pow2_ok(C, h, w, H, W, ac, S_out)) return false;
  const char* e = getenv("B200SEG_UPCELL");
  return e && e[0] == 'o';
}

long long up_fused_workspace(int N, int C, int h, int w, int H, int W, int ac) {
  if (!up_fast_ok(C, h, w, H, W, ac)) return 0;
  // corner sums PB (N,C,h+1,w+1) float4; for C > 32 (class-tiled plan) also the (N,H,W) f32 log2-sum-exp map
  return (long long)N * C * (h + 1) * (w + 1) * 4 * (long long)sizeof(float) + (C > 32 ? (long long)N * H * W * (long long)sizeof(float) : 0);
}

template <typename T>
static int up_combine_t(const void* ws, void* grad, int N, int C, int h, int w, float scale_host, const float* grad_out,
                        int use_nvalid, const uint64_t* stats, cudaStream_t st) {
  const long long tasks = (long long)N * C * ((h + kCombineRows - 1) / kCombineRows) * ((w + 30) / 31);   // one per warp
  long long blocks = (tasks + 7) / 8;
  if (blocks > kSMs * 8) blocks = kSMs * 8;
  if (blocks < 1) blocks = 1;
  launch_pdl(up_combine_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, st, reinterpret_cast<const float*>(ws),
             reinterpret_cast<T*>(grad), N * C, h, w, scale_host, grad_out, use_nvalid,
             reinterpret_cast<const unsigned long long*>(stats));
  count_launch();
  return check_launch("up_combine_kernel");
}

int up_combine_dispatch(const void* ws, void* grad, int dtype, int N, int C, int h, int w, float scale_host,
                        const float* grad_out, int use_nvalid, const uint64_t* stats, cudaStream_t st) {
  switch (dtype) {
    case B200SEG_F32: return up_combine_t<float>(ws, grad, N, C, h, w, scale_host, grad_out, use_nvalid, stats, st);
    case B200SEG_BF16: return up_combine_t<__nv_bfloat16>(ws, grad, N, C, h, w, scale_host, grad_out, use_nvalid, stats, st);
    case B200SEG_F16: return up_combine_t<__half>(ws, grad, N, C, h, w, scale_host, grad_out, use_nvalid, stats, st);
  }
  set_error("loss_fused_combine: unsupported dtype %d", dtype);
  return 1;
}

int scale_inplace_dispatch(void* x, int dtype, long long n, const float* g, cudaStream_t st) {
  long long blocks = (n + 255) / 256;
  if (blocks > kSMs * 16) blocks = kSMs * 16;
  if (blocks < 1) blocks = 1;
  switch (dtype) {
    case B200SEG_F32: scale_inplace_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((float*)x, n, g); break;
    case B200SEG_BF16: scale_inplace_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((__nv_bfloat16*)x, n, g); break;
    case B200SEG_F16: scale_inplace_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>((__half*)x, n, g); break;
    default: set_error("scale_inplace: unsupported dtype %d", dtype); return 1;
  }
  count_launch();
  return check_launch("scale_inplace_kernel");
}

template <typename T> int upcell_run(const b200seg_loss_desc* f, float* pb, int S, bool grad, cudaStream_t st);
template <typename T> int upgen_run(const b200seg_loss_desc* f, float* pb, bool grad, cudaStream_t st);

template <typename T> static int up_run(const b200seg_loss_fused_desc* d, cudaStream_t st) {
  const b200seg_loss_desc* f = &d->fwd;
  const bool grad = d->grad_logits != nullptr || d->defer_combine;
  B200SEG_REQUIRE(!grad || d->workspace != nullptr, "loss_fused: workspace is NULL");
  int S = 0;
  if (up_use_old(f->C, f->h, f->w, f->H, f->W, f->align_corners, &S)) {
    if (int e = upcell_run<T>(f, reinterpret_cast<float*>(d->workspace), S, grad, st)) return e;
  } else {
    if (int e = upgen_run<T>(f, reinterpret_cast<float*>(d->workspace), grad, st)) return e;
  }
  if (!grad || d->defer_combine) return 0;
  return up_combine_dispatch(d->workspace, d->grad_logits, f->logit_dtype, f->N, f->C, f->h, f->w, d->grad_scale_host,
                             d->grad_out, d->use_nvalid, f->stats, st);
}

int flat_fused_dispatch(const b200seg_loss_fused_desc* d, cudaStream_t st);

int up_fused_dispatch(const b200seg_loss_fused_desc* d, cudaStream_t st) {
  const b200seg_loss_desc* f = &d->fwd;
  const bool up = (f->h != f->H) || (f->w != f->W);
  if (!up) return flat_fused_dispatch(d, st);
  B200SEG_REQUIRE(up_fast_ok(f->C, f->h, f->w, f->H, f->W, f->align_corners),
                  "loss_fused: the resize-fused single pass needs H >= h, W >= w and C <= 512 "
                  "(query b200seg_loss_fused_workspace_bytes() != 0 first)");
  switch (f->logit_dtype) {
    case B200SEG_F32: return up_run<float>(d, st);
    case B200SEG_BF16: return up_run<__nv_bfloat16>(d, st);
    case B200SEG_F16: return up_run<__half>(d, st);
  }
  set_error("loss_fused: unsupported logit dtype %d", f->logit_dtype);
  return 1;
}

}  // namespace b200seg
