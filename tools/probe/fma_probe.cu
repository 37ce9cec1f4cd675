// Probe: which evaluation of ATen's bilinear expression
//   h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11)        (UpSampleBilinear2d.cu, upsample_bilinear2d_out_frame)
// (FMA contraction pattern / algebraic form) and of its source index  scale * (dst + 0.5) - 0.5  reproduces
// F.interpolate bit for bit on this GPU / torch build. Built by tools/probe/run_fma_probe.py (nvcc, in-tree); test
// infrastructure, not part of libb200seg.so.
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float comb(int v, float x, float a, float y, float b) {   // x*a + y*b
  switch (v) {
    case 0: return __fmaf_rn(x, a, __fmul_rn(y, b));
    case 1: return __fmaf_rn(y, b, __fmul_rn(x, a));
    default: return __fadd_rn(__fmul_rn(x, a), __fmul_rn(y, b));
  }
}

__device__ __forceinline__ void src_index(int idx_fma, float scale, int dst, int in, bool ac, int& i0, int& i1, float& l1) {
  float src;
  if (ac) {
    src = __fmul_rn(scale, (float)dst);
  } else {
    src = idx_fma ? __fmaf_rn(scale, __fadd_rn((float)dst, 0.5f), -0.5f) : __fadd_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), -0.5f);
    src = src < 0.f ? 0.f : src;
  }
  int i = (int)src;
  i0 = i;
  i1 = i + (i < in - 1 ? 1 : 0);
  l1 = __fsub_rn(src, (float)i);
}

// form 0: the expression as written, contraction (outer, inner)
// form 1: lerp: t = a + w1 (b - a), u = c + w1 (d - c), t + h1 (u - t)   (each a + w*(b-a) as fma when inner == 0)
// form 2: four products: (h0 w0) a + (h0 w1) b + (h1 w0) c + (h1 w1) d, summed left to right (fma chain when inner == 0)
extern "C" __global__ void probe_kernel(const float* in, float* out, int NC, int h, int w, int H, int W, int ac, float sh, float sw,
                                        int form, int outer, int inner, int idx_fma) {
  const long long total = (long long)NC * H * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(t % W), Y = (int)((t / W) % H);
    const long long nc = t / ((long long)W * H);
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(idx_fma, sh, Y, h, ac != 0, y0, y1, ly);
    src_index(idx_fma, sw, X, w, ac != 0, x0, x1, lx);
    const float h1 = ly, h0 = __fsub_rn(1.f, ly), w1 = lx, w0 = __fsub_rn(1.f, lx);
    const float* pl = in + nc * h * w;
    const float a = pl[y0 * w + x0], b = pl[y0 * w + x1], c = pl[y1 * w + x0], d = pl[y1 * w + x1];
    float r;
    if (form == 0) {
      const float Xv = comb(inner, w0, a, w1, b), Yv = comb(inner, w0, c, w1, d);
      r = comb(outer, h0, Xv, h1, Yv);
    } else if (form == 1) {
      const float tt = inner == 0 ? __fmaf_rn(w1, __fsub_rn(b, a), a) : __fadd_rn(a, __fmul_rn(w1, __fsub_rn(b, a)));
      const float uu = inner == 0 ? __fmaf_rn(w1, __fsub_rn(d, c), c) : __fadd_rn(c, __fmul_rn(w1, __fsub_rn(d, c)));
      r = outer == 0 ? __fmaf_rn(h1, __fsub_rn(uu, tt), tt) : __fadd_rn(tt, __fmul_rn(h1, __fsub_rn(uu, tt)));
    } else {
      const float p00 = __fmul_rn(h0, w0), p01 = __fmul_rn(h0, w1), p10 = __fmul_rn(h1, w0), p11 = __fmul_rn(h1, w1);
      if (inner == 0) r = __fmaf_rn(p11, d, __fmaf_rn(p10, c, __fmaf_rn(p01, b, __fmul_rn(p00, a))));
      else r = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p00, a), __fmul_rn(p01, b)), __fmul_rn(p10, c)), __fmul_rn(p11, d));
    }
    out[t] = r;
  }
}

extern "C" int probe_launch(const float* in, float* out, int NC, int h, int w, int H, int W, int ac, float sh, float sw, int form,
                            int outer, int inner, int idx_fma, void* stream) {
  probe_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(in, out, NC, h, w, H, W, ac, sh, sw, form, outer, inner, idx_fma);
  return (int)cudaGetLastError();
}
