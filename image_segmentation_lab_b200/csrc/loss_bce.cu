// Sigmoid cross-entropy with one-hot expansion fused in (the reference's use_sigmoid=True path), sm_100a.
//
// Replaces binary_cross_entropy + _expand_onehot_labels (models/losses/cross_entropy_loss.py:77-164): the
// reference materialises an (N,C,H,W) one-hot label tensor, an (N,C,H,W) valid mask and an (N,C,H,W) weight tensor
// (torch.nonzero + advanced indexing + two expands), then calls F.binary_cross_entropy_with_logits and reduces.
// Here every logit is read once; the target of element (n,c,px) is 1[label == c] (or the label itself when the
// prediction has a single channel, :126-134), the mask is (label >= 0 && label != ignore_index) (:79,:147), and
//   loss = (1 - t) x + (1 + (pos_weight_c - 1) t) * (log1p(exp(-|x|)) + max(-x, 0))        (ATen's formula)
//   grad = G * w_px * valid * ((1 - t) - (1 + (pos_weight_c - 1) t) * sigmoid(-x))
// Streaming, HBM-bound: forward C*s + L bytes per pixel, backward 2*C*s + L.
#include "common.cuh"

namespace b200seg {

struct BceParams {
  const void* logits;
  const void* labels;
  const float* pw;       // (N,HW) pixel weight or NULL
  const float* posw;     // (C) pos_weight or NULL
  float* loss_elem;      // (N,C,HW) f32 or NULL (reduction='none')
  const float* grad_out; // scalar or NULL
  const float* grad_elem;// (N,C,HW) f32 upstream gradient (reduction='none') or NULL
  void* grad;            // (N,C,HW) logit dtype or NULL
  unsigned long long* stats;   // [0] double loss sum, [1] int64 n_valid pixels
  int label_dtype;
  int N, C;
  long long HW;
  long long ignore_index;
  int single_channel;    // prediction was (N,1,H,W): the label (0/1) is the target itself
  float scale_host;      // backward: G = scale_host * (*grad_out or 1) / (use_nvalid ? n_valid*C + eps : 1)
  int use_nvalid;
  float lw;              // forward: multiplies loss_elem
};

__device__ __forceinline__ float softplus_neg(float x) {   // log(1 + exp(-x)), stable
  const float ax = fabsf(x);
  return log1pf(__expf(-ax)) + fmaxf(-x, 0.f);
}

template <typename T, int V, bool BWD>
__global__ void __launch_bounds__(256) bce_kernel(const BceParams p) {
  const int n = blockIdx.y;
  const int C = p.C;
  const long long HW = p.HW;
  const long long px0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
  float loss_acc = 0.f;
  int n_valid = 0;
  if (px0 < HW) {
    long long y[V];
    load_labels<V>(p.labels, p.label_dtype, (size_t)n * HW + px0, y);
    float wv[V];
    bool ok[V];
#pragma unroll
    for (int v = 0; v < V; ++v) wv[v] = 1.f;
    if (p.pw) {
      if constexpr (V == 8) {
        float a[4], b[4];
        load_vec<float, 4>(p.pw + (size_t)n * HW + px0, a);
        load_vec<float, 4>(p.pw + (size_t)n * HW + px0 + 4, b);
#pragma unroll
        for (int k = 0; k < 4; ++k) { wv[k] = a[k]; wv[4 + k] = b[k]; }
      } else {
        load_vec<float, V>(p.pw + (size_t)n * HW + px0, wv);
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      ok[v] = (y[v] >= 0) && (y[v] != p.ignore_index);
      n_valid += ok[v];
      if (!ok[v]) wv[v] = 0.f;
    }
    float G = 0.f;
    if constexpr (BWD) {
      G = p.scale_host * (p.grad_out ? __ldg(p.grad_out) : 1.f);
      if (p.use_nvalid) {
        const double nv = (double)(long long)p.stats[1] * (double)C;
        G = (float)((double)G / (double)(float)(nv + 1.1920928955078125e-07));
      }
    }
    const T* q = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * HW + px0;
    T* gq = BWD ? reinterpret_cast<T*>(p.grad) + (size_t)n * C * HW + px0 : nullptr;
    float* lq = (!BWD && p.loss_elem) ? p.loss_elem + (size_t)n * C * HW + px0 : nullptr;
    const float* geq = (BWD && p.grad_elem) ? p.grad_elem + (size_t)n * C * HW + px0 : nullptr;
    for (int c = 0; c < C; ++c) {
      float z[V];
      load_vec<T, V>(q, z);
      const float a1 = p.posw ? __ldg(p.posw + c) - 1.f : 0.f;
      float out[V];
      float ge[V];
      if (geq) {
        if constexpr (V == 8) {
          float a[4], b[4];
          load_vec<float, 4>(geq, a);
          load_vec<float, 4>(geq + 4, b);
#pragma unroll
          for (int k = 0; k < 4; ++k) { ge[k] = a[k]; ge[4 + k] = b[k]; }
        } else {
          load_vec<float, V>(geq, ge);
        }
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float t = p.single_channel ? (float)y[v] : ((long long)c == y[v] ? 1.f : 0.f);
        const float a = fmaf(a1, t, 1.f);
        if constexpr (BWD) {
          const float sg = 1.f / (1.f + __expf(z[v]));   // sigmoid(-x)
          float g = G * wv[v] * ((1.f - t) - a * sg);
          if (geq) g *= ge[v];
          out[v] = g;
        } else {
          const float l = wv[v] * fmaf(1.f - t, z[v], a * softplus_neg(z[v]));
          loss_acc += l;
          out[v] = l * p.lw;
        }
      }
      if constexpr (BWD) {
        store_vec<T, V>(gq, out);
        gq += HW;
        if (geq) geq += HW;
      } else if (lq) {
        if constexpr (V == 8) {
          float a[4] = {out[0], out[1], out[2], out[3]}, b[4] = {out[4], out[5], out[6], out[7]};
          store_vec<float, 4>(lq, a);
          store_vec<float, 4>(lq + 4, b);
        } else {
          store_vec<float, V>(lq, out);
        }
        lq += HW;
      }
      q += HW;
    }
  }
  if constexpr (!BWD) {
    __shared__ float s_l[32];
    __shared__ int s_n[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    loss_acc = warp_sum(loss_acc);
    n_valid = __reduce_add_sync(0xffffffffu, n_valid);
    if (lane == 0) { s_l[warp] = loss_acc; s_n[warp] = n_valid; }
    __syncthreads();
    if (warp == 0) {
      double l = lane < 8 ? (double)s_l[lane] : 0.0;
      l = warp_sum(l);
      const int nv = __reduce_add_sync(0xffffffffu, lane < 8 ? s_n[lane] : 0);
      if (lane == 0) {
        atomicAdd(reinterpret_cast<double*>(p.stats), l);
        atomicAdd(p.stats + 1, (unsigned long long)nv);
      }
    }
  }
}

template <typename T, bool BWD> static int launch_bce(const BceParams& p, bool vec, cudaStream_t st) {
  constexpr int VV = 16 / (int)sizeof(T);
  if (vec) {
    dim3 grid((unsigned)((p.HW / VV + 255) / 256), p.N);
    bce_kernel<T, VV, BWD><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid((unsigned)((p.HW + 255) / 256), p.N);
    bce_kernel<T, 1, BWD><<<grid, 256, 0, st>>>(p);
  }
  count_launch();
  return check_launch("bce_kernel");
}

}  // namespace b200seg

using namespace b200seg;

static int bce_common(const b200seg_bce_desc* d, bool bwd, void* stream) {
  B200SEG_REQUIRE(d != nullptr, "bce: NULL descriptor");
  B200SEG_REQUIRE(d->N >= 0 && d->N <= 65535 && d->C >= 1 && d->HW >= 1, "bce: bad shape N=%d C=%d HW=%lld", d->N, d->C,
                  (long long)d->HW);
  B200SEG_REQUIRE(d->label_dtype >= B200SEG_L_U8 && d->label_dtype <= B200SEG_L_F64, "bce: unsupported label dtype");
  B200SEG_REQUIRE(d->stats != nullptr, "bce: NULL stats");
  B200SEG_REQUIRE(!d->single_channel || d->C == 1, "bce: single_channel needs C == 1");
  cudaStream_t st = (cudaStream_t)stream;
  if (!bwd) B200SEG_CUDA(cudaMemsetAsync(d->stats, 0, 2 * sizeof(uint64_t), st));
  if (d->N == 0) return 0;
  B200SEG_REQUIRE(d->logits && d->labels, "bce: NULL logits/labels");
  B200SEG_REQUIRE(!bwd || d->grad_logits, "bce_bwd: NULL grad_logits");
  BceParams p = {};
  p.logits = d->logits; p.labels = d->labels; p.pw = d->pixel_weight; p.posw = d->pos_weight;
  p.loss_elem = d->loss_elem; p.grad_out = d->grad_out; p.grad_elem = d->grad_elem; p.grad = d->grad_logits;
  p.stats = reinterpret_cast<unsigned long long*>(d->stats);
  p.label_dtype = d->label_dtype; p.N = d->N; p.C = d->C; p.HW = d->HW;
  p.ignore_index = d->ignore_index; p.single_channel = d->single_channel;
  p.scale_host = d->grad_scale_host; p.use_nvalid = d->use_nvalid; p.lw = d->loss_weight;
  const int VV = 16 / logit_bytes(d->logit_dtype);
  const bool vec = (d->HW % VV == 0) && aligned16(d->logits) && aligned16(d->labels) && (!p.pw || aligned16(p.pw)) &&
                   (!p.loss_elem || aligned16(p.loss_elem)) && (!p.grad || aligned16(p.grad)) &&
                   (!p.grad_elem || aligned16(p.grad_elem));
  switch (d->logit_dtype) {
    case B200SEG_F32: return bwd ? launch_bce<float, true>(p, vec, st) : launch_bce<float, false>(p, vec, st);
    case B200SEG_BF16:
      return bwd ? launch_bce<__nv_bfloat16, true>(p, vec, st) : launch_bce<__nv_bfloat16, false>(p, vec, st);
    case B200SEG_F16: return bwd ? launch_bce<__half, true>(p, vec, st) : launch_bce<__half, false>(p, vec, st);
  }
  set_error("bce: unsupported logit dtype %d", d->logit_dtype);
  return 1;
}

extern "C" int b200seg_bce_fwd(const b200seg_bce_desc* d, void* stream) { return bce_common(d, false, stream); }
extern "C" int b200seg_bce_bwd(const b200seg_bce_desc* d, void* stream) { return bce_common(d, true, stream); }
