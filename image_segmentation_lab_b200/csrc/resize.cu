// Stand-alone resize kernels (reference utils/ops.py:7-26 -> F.interpolate), sm_100a.
//
// Forward: bit-identical to ATen's upsample_bilinear2d (contraction pinned, see common.cuh). One thread per V
// consecutive output pixels of one (n,c) plane; the four taps come
// through the read-only path (each input element is re-read by ~scale^2 neighbours, all L1 hits),
// the output is written once with 128-bit streaming stores. HBM-bound on the output write.
// Backward: deterministic GATHER form of the transpose — one thread per input element sums the
// output-gradient pixels whose taps touch it, in a fixed order (ATen uses an atomicAdd scatter,
// which is non-deterministic).
#include "common.cuh"

namespace b200seg {

template <typename T, int V>
__global__ void __launch_bounds__(256) resize_bilinear_fwd_kernel(const T* __restrict__ in, T* __restrict__ out, int NC,
                                                                  int h, int w, int H, int W, float sh, float sw,
                                                                  int ac) {
  const long long per_row = W / V;
  const long long total = (long long)NC * H * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xv = (int)(i % per_row);
    const long long r = i / per_row;
    const int Y = (int)(r % H);
    const long long nc = r / H;
    int y0, y1;
    float ly;
    resize_src(sh, Y, h, ac != 0, y0, y1, ly);
    const float h1 = ly, h0 = __fsub_rn(1.f, ly);
    const T* r0 = in + ((size_t)nc * h + y0) * w;
    const T* r1 = in + ((size_t)nc * h + y1) * w;
    float o[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      int x0, x1;
      float lx;
      resize_src(sw, xv * V + v, w, ac != 0, x0, x1, lx);
      const float w1 = lx, w0 = __fsub_rn(1.f, lx);
      // ATen's operations, contraction pinned (common.cuh): bit-identical to F.interpolate
      o[v] = aten_bilerp(h0, h1, w0, w1, to_float<T>(r0[x0]), to_float<T>(r0[x1]), to_float<T>(r1[x0]), to_float<T>(r1[x1]));
    }
    store_vec<T, V>(out + ((size_t)nc * H + Y) * W + (size_t)xv * V, o);
  }
}

// range of output indices whose taps can touch input index i (conservative; exact taps re-checked)
__device__ __forceinline__ void touch_range(float scale, int i, int out, bool ac, int& lo, int& hi) {
  if (scale <= 0.f) { lo = 0; hi = out - 1; return; }
  float a, b;
  if (ac) { a = ((float)i - 1.f) / scale; b = ((float)i + 1.f) / scale; }
  else { a = ((float)i - 0.5f) / scale - 0.5f; b = ((float)i + 1.5f) / scale - 0.5f; }
  lo = (int)floorf(a) - 1;
  hi = (int)ceilf(b) + 1;
  lo = lo < 0 ? 0 : lo;
  hi = hi > out - 1 ? out - 1 : hi;
}

template <typename T>
__global__ void __launch_bounds__(256) resize_bilinear_bwd_kernel(const T* __restrict__ go, T* __restrict__ gi, int NC,
                                                                  int h, int w, int H, int W, float sh, float sw,
                                                                  int ac) {
  const long long total = (long long)NC * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const long long r = i / w;
    const int y = (int)(r % h);
    const long long nc = r / h;
    int Ylo, Yhi, Xlo, Xhi;
    touch_range(sh, y, H, ac != 0, Ylo, Yhi);
    touch_range(sw, x, W, ac != 0, Xlo, Xhi);
    float acc = 0.f;
    for (int Y = Ylo; Y <= Yhi; ++Y) {
      int y0, y1;
      float ly;
      resize_src(sh, Y, h, ac != 0, y0, y1, ly);
      const float wy = (y0 == y ? 1.f - ly : 0.f) + (y1 == y ? ly : 0.f);
      if (wy == 0.f) continue;
      const T* row = go + ((size_t)nc * H + Y) * W;
      float racc = 0.f;
      for (int X = Xlo; X <= Xhi; ++X) {
        int x0, x1;
        float lx;
        resize_src(sw, X, w, ac != 0, x0, x1, lx);
        const float wx = (x0 == x ? 1.f - lx : 0.f) + (x1 == x ? lx : 0.f);
        if (wx != 0.f) racc = fmaf(wx, to_float<T>(row[X]), racc);
      }
      acc = fmaf(wy, racc, acc);
    }
    gi[i] = from_float<T>(acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) resize_nearest_fwd_kernel(const T* __restrict__ in, T* __restrict__ out, int NC,
                                                                 int h, int w, int H, int W, float sh, float sw) {
  const long long total = (long long)NC * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    const long long r = i / W;
    const int Y = (int)(r % H);
    const long long nc = r / H;
    int y = (int)floorf((float)Y * sh), x = (int)floorf((float)X * sw);
    y = y < h - 1 ? y : h - 1;
    x = x < w - 1 ? x : w - 1;
    out[i] = in[((size_t)nc * h + y) * w + x];
  }
}

// first output index in [0, out] whose nearest source min(floor(dst * scale), in - 1) is >= i
__device__ __forceinline__ int nearest_start(float scale, int i, int in, int out) {
  if (i <= 0) return 0;
  if (i > in - 1 || !(scale > 0.f)) return out;
  int d = (int)fminf(fmaxf(ceilf((float)i / scale), 0.f), (float)out);
  auto src = [&](int dst) { const int v = (int)floorf((float)dst * scale); return v < in - 1 ? v : in - 1; };
  while (d > 0 && src(d - 1) >= i) --d;
  while (d < out && src(d) < i) ++d;
  return d;
}

// Backward of the nearest resize, deterministic gather: every input element sums the output gradients that copied it.
template <typename T>
__global__ void __launch_bounds__(256) resize_nearest_bwd_kernel(const T* __restrict__ go, T* __restrict__ gi, int NC,
                                                                 int h, int w, int H, int W, float sh, float sw) {
  const long long total = (long long)NC * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const long long r = i / w;
    const int y = (int)(r % h);
    const long long nc = r / h;
    const int Y0 = nearest_start(sh, y, h, H), Y1 = nearest_start(sh, y + 1, h, H);
    const int X0 = nearest_start(sw, x, w, W), X1 = nearest_start(sw, x + 1, w, W);
    float acc = 0.f;
    for (int Y = Y0; Y < Y1; ++Y) {
      const T* row = go + ((size_t)nc * H + Y) * W;
      for (int X = X0; X < X1; ++X) acc += to_float<T>(row[X]);
    }
    gi[i] = from_float<T>(acc);
  }
}

static int grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)kSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ATen's compute_scales_value: an explicit scale_factor replaces in / out by (float)(1.0 / scale_factor) — align_corners
// ignores it. scale_override <= 0 means "from the sizes".
static float pick_scale(int in, int out, int ac, float scale_override) {
  if (!ac && scale_override > 0.f) return scale_override;
  return resize_scale(in, out, ac != 0);
}

template <typename T>
static int resize_fwd_t(const void* in, void* out, int NC, int h, int w, int H, int W, int ac, float so_h, float so_w, cudaStream_t st) {
  const float sh = pick_scale(h, H, ac, so_h), sw = pick_scale(w, W, ac, so_w);
  constexpr int VV = 16 / (int)sizeof(T);
  if (W % VV == 0 && aligned16(out)) {
    resize_bilinear_fwd_kernel<T, VV><<<grid_for((long long)NC * H * (W / VV)), 256, 0, st>>>(
        (const T*)in, (T*)out, NC, h, w, H, W, sh, sw, ac);
  } else {
    resize_bilinear_fwd_kernel<T, 1><<<grid_for((long long)NC * H * W), 256, 0, st>>>((const T*)in, (T*)out, NC, h, w,
                                                                                       H, W, sh, sw, ac);
  }
  count_launch();
  return check_launch("resize_bilinear_fwd_kernel");
}

template <typename T>
static int resize_bwd_t(const void* go, void* gi, int NC, int h, int w, int H, int W, int ac, float so_h, float so_w, cudaStream_t st) {
  const float sh = pick_scale(h, H, ac, so_h), sw = pick_scale(w, W, ac, so_w);
  resize_bilinear_bwd_kernel<T><<<grid_for((long long)NC * h * w), 256, 0, st>>>((const T*)go, (T*)gi, NC, h, w, H, W,
                                                                                 sh, sw, ac);
  count_launch();
  return check_launch("resize_bilinear_bwd_kernel");
}

template <typename T>
static int resize_nearest_t(const void* in, void* out, int NC, int h, int w, int H, int W, float so_h, float so_w, bool bwd,
                            cudaStream_t st) {
  // ATen nearest: scale = in/out in fp32 (UpSample.h compute_scales_value), src = min(floor(dst*scale), in-1)
  const float sh = pick_scale(h, H, 0, so_h), sw = pick_scale(w, W, 0, so_w);
  if (bwd) {
    resize_nearest_bwd_kernel<T><<<grid_for((long long)NC * h * w), 256, 0, st>>>((const T*)in, (T*)out, NC, h, w, H, W, sh, sw);
    count_launch();
    return check_launch("resize_nearest_bwd_kernel");
  }
  resize_nearest_fwd_kernel<T><<<grid_for((long long)NC * H * W), 256, 0, st>>>((const T*)in, (T*)out, NC, h, w, H, W,
                                                                                sh, sw);
  count_launch();
  return check_launch("resize_nearest_fwd_kernel");
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int b200seg_resize_bilinear_fwd(const void* in, void* out, int32_t dtype, int32_t NC, int32_t h, int32_t w,
                                           int32_t H, int32_t W, int32_t align_corners, float scale_h, float scale_w,
                                           void* stream) {
  B200SEG_REQUIRE(in && out, "resize: NULL tensor");
  B200SEG_REQUIRE(NC >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "resize: bad shape");
  if (NC == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (h == H && w == W) {  // F.interpolate returns a copy
    B200SEG_CUDA(cudaMemcpyAsync(out, in, (size_t)NC * h * w * logit_bytes(dtype), cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  switch (dtype) {
    case B200SEG_F32: return resize_fwd_t<float>(in, out, NC, h, w, H, W, align_corners, scale_h, scale_w, st);
    case B200SEG_BF16: return resize_fwd_t<__nv_bfloat16>(in, out, NC, h, w, H, W, align_corners, scale_h, scale_w, st);
    case B200SEG_F16: return resize_fwd_t<__half>(in, out, NC, h, w, H, W, align_corners, scale_h, scale_w, st);
  }
  set_error("resize: unsupported dtype %d", dtype);
  return 1;
}

extern "C" int b200seg_resize_bilinear_bwd(const void* go, void* gi, int32_t dtype, int32_t NC, int32_t h, int32_t w,
                                           int32_t H, int32_t W, int32_t align_corners, float scale_h, float scale_w,
                                           void* stream) {
  B200SEG_REQUIRE(go && gi, "resize_bwd: NULL tensor");
  B200SEG_REQUIRE(NC >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "resize_bwd: bad shape");
  if (NC == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (h == H && w == W) {
    B200SEG_CUDA(cudaMemcpyAsync(gi, go, (size_t)NC * h * w * logit_bytes(dtype), cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  switch (dtype) {
    case B200SEG_F32: return resize_bwd_t<float>(go, gi, NC, h, w, H, W, align_corners, scale_h, scale_w, st);
    case B200SEG_BF16: return resize_bwd_t<__nv_bfloat16>(go, gi, NC, h, w, H, W, align_corners, scale_h, scale_w, st);
    case B200SEG_F16: return resize_bwd_t<__half>(go, gi, NC, h, w, H, W, align_corners, scale_h, scale_w, st);
  }
  set_error("resize_bwd: unsupported dtype %d", dtype);
  return 1;
}

static int nearest_dispatch(const void* a, void* b, int dtype, int NC, int h, int w, int H, int W, float so_h, float so_w, bool bwd,
                            cudaStream_t st) {
  switch (dtype) {
    case B200SEG_F32: return resize_nearest_t<float>(a, b, NC, h, w, H, W, so_h, so_w, bwd, st);
    case B200SEG_BF16: return resize_nearest_t<__nv_bfloat16>(a, b, NC, h, w, H, W, so_h, so_w, bwd, st);
    case B200SEG_F16: return resize_nearest_t<__half>(a, b, NC, h, w, H, W, so_h, so_w, bwd, st);
  }
  set_error("resize_nearest: unsupported dtype %d", dtype);
  return 1;
}

extern "C" int b200seg_resize_nearest_fwd(const void* in, void* out, int32_t dtype, int32_t NC, int32_t h, int32_t w,
                                          int32_t H, int32_t W, float scale_h, float scale_w, void* stream) {
  B200SEG_REQUIRE(in && out, "resize_nearest: NULL tensor");
  B200SEG_REQUIRE(NC >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "resize_nearest: bad shape");
  if (NC == 0) return 0;
  return nearest_dispatch(in, out, dtype, NC, h, w, H, W, scale_h, scale_w, false, (cudaStream_t)stream);
}

extern "C" int b200seg_resize_nearest_bwd(const void* grad_out, void* grad_in, int32_t dtype, int32_t NC, int32_t h, int32_t w,
                                          int32_t H, int32_t W, float scale_h, float scale_w, void* stream) {
  B200SEG_REQUIRE(grad_out && grad_in, "resize_nearest_bwd: NULL tensor");
  B200SEG_REQUIRE(NC >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "resize_nearest_bwd: bad shape");
  if (NC == 0) return 0;
  return nearest_dispatch(grad_out, grad_in, dtype, NC, h, w, H, W, scale_h, scale_w, true, (cudaStream_t)stream);
}
