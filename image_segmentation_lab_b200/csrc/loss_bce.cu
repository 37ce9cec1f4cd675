// Sigmoid cross-entropy with one-hot expansion fused in (the reference's use_sigmoid=True path), sm_100a.
//
// Replaces binary_cross_entropy + _expand_onehot_labels (models/losses/cross_entropy_loss.py:77-164): the
// reference materialises an (N,C,H,W) one-hot label tensor, an (N,C,H,W) valid mask and an (N,C,H,W) weight tensor
// (torch.nonzero + advanced indexing + two expands), then calls F.binary_cross_entropy_with_logits and reduces.
// Here every logit is read once; the target of element (n,c,px) is 1[label == c] (or the label itself when the
// prediction has a single channel, :126-134), the mask is (label >= 0 && label != ignore_index) (:79,:147), and
//   loss = (1 - t) x + (1 + (pos_weight_c - 1) t) * (log1p(exp(-|x|)) + max(-x, 0))        (ATen's formula)
//   grad = G * w_px * valid * ((1 - t) - (1 + (pos_weight_c - 1) t) * sigmoid(-x))
// Streaming, HBM-bound. Three modes of one persistent kernel (grid = a few CTAs per SM walking 256 x V pixel blocks):
//   forward   C*s + L bytes per pixel: loss sums (fp64) + valid-pixel count; the last CTA to finish writes the reduced
//             scalar (mean / sum / avg_factor / avg_non_ignore), so that the host issues no arithmetic of its own;
//   backward  2*C*s + L: the gradient from the logits, the labels and the upstream gradient on the device;
//   fused     2*C*s + L for forward AND backward: when the denominator is known before the launch (every reduction but
//             avg_non_ignore) the gradient is written in the forward pass with the upstream gradient taken as 1 and
//             rescaled later only if it is not (b200seg_scale_inplace returns at once when it is).
// exp(-|x|) by one MUFU.EX2, 1/(1+e) by one MUFU.RCP (the sigmoid), log1p(e) = 2 atanh(e / (2 + e)) by a second
// reciprocal and a 6-term odd series (|z| <= 1/3: truncation 1.4e-8 relative) instead of log1pf's ~30 instructions.
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
  unsigned long long* stats;   // [0] double loss sum, [1] int64 n_valid pixels, [2] CTAs done, [3] top-1 hits, [4] pixels counted
  float* out;            // forward: reduced scalar or NULL
  float* acc_out;        // forward: top-1 accuracy (accuracy.py:41-60) or NULL
  long long acc_ignore;  // accuracy: pixels with this label are not counted (acc_has_ignore)
  int acc_has_ignore;
  int label_dtype;
  int N, C;
  long long HW;
  long long ignore_index;
  int single_channel;    // prediction was (N,1,H,W): the label (0/1) is the target itself
  float scale_host;      // gradient: G = scale_host * (*grad_out or 1) / (use_nvalid ? n_valid*C + eps : 1)
  int use_nvalid;
  float lw;              // forward: multiplies loss_elem
  float out_scale;       // forward: out = out_scale * sum  [/ (n_valid*C + eps) when use_nvalid]
  int blocks_per_img;    // 256 x V pixel blocks per image
  long long total_blocks;
};

constexpr int kBceFwd = 0, kBceBwd = 1, kBceFused = 2;

// log(1 + exp(-x)) and sigmoid(-x) = 1 / (1 + exp(x)). log1p(e) = lg2(1 + e) ln 2 on the MUFU: 1 + e is in [1, 2], so the
// rounding of the sum and lg2.approx each cost ~1e-7 ABSOLUTE (on per-element losses of order 0.1 - 1; the per-element
// output of reduction='none' is compared in the max norm, the reduced loss at 1e-5) — 3 instructions instead of ~30
__device__ __forceinline__ void softplus_sigmoid_neg(float x, float& sp, float& sg) {
  const float e = ex2(-fabsf(x) * kLog2e);            // exp(-|x|) in [0, 1]
  const float u = 1.f + e;
  const float r = fast_rcp(u);
  sg = x >= 0.f ? e * r : r;
  float l2;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
  sp = fmaf(l2, 0.693147180559945f, fmaxf(-x, 0.f));
}

template <typename T, int V, int MODE>
__global__ void __launch_bounds__(256, V == 8 ? 4 : 6) bce_kernel(const BceParams p) {
  constexpr bool kLoss = MODE != kBceBwd, kGrad = MODE != kBceFwd;
  const int C = p.C;
  const long long HW = p.HW;
  float loss_acc = 0.f;
  int n_valid = 0, n_correct = 0, n_acc = 0;
  const bool want_acc = kLoss && p.acc_out != nullptr;
  float G = 0.f;
  if constexpr (kGrad) {
    G = p.scale_host * ((MODE == kBceBwd && p.grad_out) ? __ldg(p.grad_out) : 1.f);
    if (MODE == kBceBwd && p.use_nvalid) {
      const double nv = (double)(long long)p.stats[1] * (double)C;
      G = (float)((double)G / (double)(float)(nv + 1.1920928955078125e-07));
    }
  }
  const int ign32 = ((unsigned long long)p.ignore_index < 0x40000000ull) ? (int)p.ignore_index : -2;
  const int acc_ign32 = ((unsigned long long)p.acc_ignore < 0x40000000ull) ? (int)p.acc_ignore : -2;
  for (unsigned t = blockIdx.x; t < (unsigned)p.total_blocks; t += gridDim.x) {      // total_blocks < 2^31 (checked on the host)
    const unsigned n = t / (unsigned)p.blocks_per_img;
    const long long px0 = (long long)((t - n * (unsigned)p.blocks_per_img) * 256u + threadIdx.x) * V;
    if (px0 >= HW) continue;
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
    // labels as 32-bit class ids: -1 = not a valid pixel (negative, ignore_index, or beyond 2^30)
    int yi[V], yr[V];                                     // yr: the label as the accuracy sees it (-1: matches no class)
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int y32 = ((unsigned long long)y[v] < 0x40000000ull) ? (int)y[v] : -1;
      yr[v] = y32;
      yi[v] = (y32 == ign32) ? -1 : y32;
      ok[v] = yi[v] >= 0;
      n_valid += ok[v];
      if (!ok[v]) wv[v] = 0.f;
    }
    const bool single = p.single_channel != 0;
    float zmax[V];
    int amax[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { zmax[v] = -INFINITY; amax[v] = 0; }
    const T* q = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * HW + px0;
    T* gq = kGrad ? reinterpret_cast<T*>(p.grad) + (size_t)n * C * HW + px0 : nullptr;
    float* lq = (MODE == kBceFwd && p.loss_elem) ? p.loss_elem + (size_t)n * C * HW + px0 : nullptr;
    const float* geq = (MODE == kBceBwd && p.grad_elem) ? p.grad_elem + (size_t)n * C * HW + px0 : nullptr;
    for (int c = 0; c < C; ++c) {
      float z[V];
      load_vec<T, V>(q, z);
      const float a1 = p.posw ? __ldg(p.posw + c) - 1.f : 0.f;
      float gout[V], lout[V];
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
        const float tt = single ? (float)yi[v] : (c == yi[v] ? 1.f : 0.f);   // an invalid pixel has weight 0
        const float a = fmaf(a1, tt, 1.f);
        float sp, sg;
        softplus_sigmoid_neg(z[v], sp, sg);
        if constexpr (kLoss) {                            // arg-max of the logits = arg-max of the sigmoids; lowest index on ties
          const bool gt = z[v] > zmax[v];
          zmax[v] = gt ? z[v] : zmax[v];
          amax[v] = gt ? c : amax[v];
        }
        if constexpr (kGrad) {
          float g = G * wv[v] * ((1.f - tt) - a * sg);
          if (geq) g *= ge[v];
          gout[v] = g;
        }
        if constexpr (kLoss) {
          const float l = wv[v] * fmaf(1.f - tt, z[v], a * sp);
          loss_acc += l;
          lout[v] = l * p.lw;
        }
      }
      if constexpr (kGrad) {
        store_vec<T, V>(gq, gout);
        gq += HW;
        if (geq) geq += HW;
      }
      if (lq) {
        if constexpr (V == 8) {
          float a[4] = {lout[0], lout[1], lout[2], lout[3]}, b[4] = {lout[4], lout[5], lout[6], lout[7]};
          store_vec<float, 4>(lq, a);
          store_vec<float, 4>(lq + 4, b);
        } else {
          store_vec<float, V>(lq, lout);
        }
        lq += HW;
      }
      q += HW;
    }
    if (want_acc) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        // (a negative acc_ignore label cannot be told from other negative labels here: both are "counted, never a hit",
        // while the reference drops the former from the denominator — labels are non-negative in every dataset of the path)
        const bool counted = !p.acc_has_ignore || yr[v] != acc_ign32;
        n_acc += counted;
        n_correct += counted && yr[v] == amax[v];
      }
    }
  }
  if constexpr (kLoss) {
    __shared__ float s_l[32];
    __shared__ int s_n[32], s_c[32], s_a[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    loss_acc = warp_sum(loss_acc);
    n_valid = __reduce_add_sync(0xffffffffu, n_valid);
    n_correct = __reduce_add_sync(0xffffffffu, n_correct);
    n_acc = __reduce_add_sync(0xffffffffu, n_acc);
    if (lane == 0) { s_l[warp] = loss_acc; s_n[warp] = n_valid; s_c[warp] = n_correct; s_a[warp] = n_acc; }
    __syncthreads();
    if (warp == 0) {
      double l = lane < 8 ? (double)s_l[lane] : 0.0;
      l = warp_sum(l);
      const int nv = __reduce_add_sync(0xffffffffu, lane < 8 ? s_n[lane] : 0);
      const int nc = __reduce_add_sync(0xffffffffu, lane < 8 ? s_c[lane] : 0);
      const int na = __reduce_add_sync(0xffffffffu, lane < 8 ? s_a[lane] : 0);
      if (lane == 0) {
        atomicAdd(reinterpret_cast<double*>(p.stats), l);
        atomicAdd(p.stats + 1, (unsigned long long)nv);
        if (want_acc) {
          atomicAdd(p.stats + 3, (unsigned long long)nc);
          atomicAdd(p.stats + 4, (unsigned long long)na);
        }
        __threadfence();
        const unsigned long long done = atomicAdd(p.stats + 2, 1ull);
        if (done == (unsigned long long)gridDim.x - 1) {                     // last CTA: the reduced scalars
          __threadfence();
          constexpr double eps = 1.1920928955078125e-07;
          if (p.out) {
            const double tot = __longlong_as_double((long long)atomicAdd(p.stats, 0ull));
            double r = (double)p.out_scale * tot;
            if (p.use_nvalid) {
              const double nvt = (double)(long long)atomicAdd(p.stats + 1, 0ull) * (double)C;
              r /= (double)(float)(nvt + eps);
            }
            *p.out = (float)r;
          }
          if (want_acc) {   // accuracy.py:55-60: (correct.float().sum() + eps) * (100.0 / (n + eps)), evaluated in fp32
            const double nct = (double)(long long)atomicAdd(p.stats + 3, 0ull), nat = (double)(long long)atomicAdd(p.stats + 4, 0ull);
            *p.acc_out = ((float)nct + (float)eps) * (float)(100.0 / (nat + eps));
          }
        }
      }
    }
  }
}

template <typename T, int MODE> static int launch_bce(BceParams& p, bool vec, cudaStream_t st) {
  constexpr int VV = 16 / (int)sizeof(T);
  const int v = vec ? VV : 1;
  p.blocks_per_img = (int)((p.HW + 256ll * v - 1) / (256ll * v));
  p.total_blocks = (long long)p.N * p.blocks_per_img;
  if (p.total_blocks >= (1ll << 31)) { set_error("bce: %lld pixel blocks exceed 2^31-1", p.total_blocks); return 1; }
  // persistent grid = exactly the CTAs that are resident at once (a partial second wave would leave SMs idle)
  static int occ_v = 0, occ_1 = 0;
  int& occ = vec ? occ_v : occ_1;
  if (occ == 0) {
    int o = 0;
    cudaError_t e = vec ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, bce_kernel<T, VV, MODE>, 256, 0)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, bce_kernel<T, 1, MODE>, 256, 0);
    occ = (e == cudaSuccess && o > 0) ? o : 4;
  }
  long long grid = p.total_blocks < (long long)kSMs * occ ? p.total_blocks : (long long)kSMs * occ;
  if (grid < 1) grid = 1;
  if (vec) bce_kernel<T, VV, MODE><<<(unsigned)grid, 256, 0, st>>>(p);
  else bce_kernel<T, 1, MODE><<<(unsigned)grid, 256, 0, st>>>(p);
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
  const bool fused = !bwd && d->grad_logits != nullptr;        // forward that also writes the gradient
  B200SEG_REQUIRE(!fused || (!d->use_nvalid && !d->loss_elem),
                  "bce_fwd: the single-pass gradient needs a denominator known before the launch (no avg_non_ignore, no "
                  "per-element loss)");
  if (!bwd) B200SEG_CUDA(cudaMemsetAsync(d->stats, 0, 8 * sizeof(uint64_t), st));
  if (d->N == 0) {
    if (!bwd && d->out) B200SEG_CUDA(cudaMemsetAsync(d->out, 0, sizeof(float), st));
    if (!bwd && d->acc_out) {   // (0 + eps) * (100 / (0 + eps)) = 100 in fp32, as the reference's accuracy of an empty batch
      const float hundred = 100.f;
      B200SEG_CUDA(cudaMemcpyAsync(d->acc_out, &hundred, sizeof(float), cudaMemcpyHostToDevice, st));
    }
    return 0;
  }
  B200SEG_REQUIRE(d->logits && d->labels, "bce: NULL logits/labels");
  B200SEG_REQUIRE(!bwd || d->grad_logits, "bce_bwd: NULL grad_logits");
  BceParams p = {};
  p.logits = d->logits; p.labels = d->labels; p.pw = d->pixel_weight; p.posw = d->pos_weight;
  p.loss_elem = d->loss_elem; p.grad_out = d->grad_out; p.grad_elem = d->grad_elem; p.grad = d->grad_logits;
  p.stats = reinterpret_cast<unsigned long long*>(d->stats);
  p.out = d->out; p.out_scale = d->out_scale_host;
  p.acc_out = bwd ? nullptr : d->acc_out; p.acc_ignore = d->acc_ignore_index; p.acc_has_ignore = d->acc_has_ignore;
  p.label_dtype = d->label_dtype; p.N = d->N; p.C = d->C; p.HW = d->HW;
  p.ignore_index = d->ignore_index; p.single_channel = d->single_channel;
  p.scale_host = d->grad_scale_host; p.use_nvalid = d->use_nvalid; p.lw = d->loss_weight;
  const int VV = 16 / logit_bytes(d->logit_dtype);
  const bool vec = (d->HW % VV == 0) && aligned16(d->logits) && aligned16(d->labels) && (!p.pw || aligned16(p.pw)) &&
                   (!p.loss_elem || aligned16(p.loss_elem)) && (!p.grad || aligned16(p.grad)) &&
                   (!p.grad_elem || aligned16(p.grad_elem));
#define BCE_DISPATCH(T)                                                          \
  return bwd ? launch_bce<T, kBceBwd>(p, vec, st)                                \
             : (fused ? launch_bce<T, kBceFused>(p, vec, st) : launch_bce<T, kBceFwd>(p, vec, st))
  switch (d->logit_dtype) {
    case B200SEG_F32: BCE_DISPATCH(float);
    case B200SEG_BF16: BCE_DISPATCH(__nv_bfloat16);
    case B200SEG_F16: BCE_DISPATCH(__half);
  }
#undef BCE_DISPATCH
  set_error("bce: unsupported logit dtype %d", d->logit_dtype);
  return 1;
}

extern "C" int b200seg_bce_fwd(const b200seg_bce_desc* d, void* stream) { return bce_common(d, false, stream); }
extern "C" int b200seg_bce_bwd(const b200seg_bce_desc* d, void* stream) { return bce_common(d, true, stream); }
