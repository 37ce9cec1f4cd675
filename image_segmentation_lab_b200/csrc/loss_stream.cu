// Streaming softmax cross-entropy (+ top-1 accuracy) forward / backward for sm_100a.
//
// Replaces, in one pass over the logits each way (reference file:line):
//   F.interpolate bilinear            utils/ops.py:26            (general-ratio variant, UP=true)
//   F.cross_entropy(reduction='none') models/losses/cross_entropy_loss.py:56-61
//   weight / reduce                   models/losses/utils.py:48-80
//   accuracy top-1                    models/losses/accuracy.py:41-60
//
// Data layout: logits NCHW, so for a fixed class the pixels of a row are contiguous. A thread owns
// V consecutive pixels (V*sizeof(T) = 16 bytes) and walks the class dimension with stride H*W;
// a warp therefore issues one fully coalesced 512-byte request per class. Classes are consumed in
// register chunks of CH with an online (running max / rescaled sum) soft-max, so every logit is
// read from HBM exactly once and never written back.
//
// Roofline: HBM. Algorithmic bytes per pixel: fwd C*s + L (+4 lse), bwd 2*C*s + L + 4.
#include "common.cuh"

namespace b200seg {

struct CeFwdParams {
  const void* logits;
  const void* labels;
  const float* pw;
  const float* cw;
  float* lse;
  float* loss_px;
  unsigned long long* stats;
  int label_dtype;
  int N, C, h, w, H, W;
  int align_corners;
  int flags;
  long long ignore_index;
  int acc_has_ignore;
  long long acc_ignore;
  float lw;
  float sh, sw;
  double* dice_part;        // (N,C,3) or NULL: also accumulate the one-hot Dice sums [sum p_y*valid, -, count]
  long long dice_ignore;
  int tversky;              // B200SEG_MODE_TVERSKY: count only valid pixels, store lse = +inf for ignored ones
};

template <int V> __device__ __forceinline__ void load_f32(const float* p, float (&o)[V]) {
  if constexpr (V == 8) {
    float a[4], b[4];
    load_vec<float, 4>(p, a);
    load_vec<float, 4>(p + 4, b);
#pragma unroll
    for (int k = 0; k < 4; ++k) { o[k] = a[k]; o[4 + k] = b[k]; }
  } else {
    load_vec<float, V>(p, o);
  }
}
template <int V> __device__ __forceinline__ void store_f32(float* p, const float (&v)[V]) {
  if constexpr (V == 8) {
    float a[4] = {v[0], v[1], v[2], v[3]}, b[4] = {v[4], v[5], v[6], v[7]};
    store_vec<float, 4>(p, a);
    store_vec<float, 4>(p + 4, b);
  } else {
    store_vec<float, V>(p, v);
  }
}

// Bilinear taps of one output pixel (UP variant).
struct Taps {
  int o00, o01, o10, o11;
  float w00, w01, w10, w11;
  float h0, h1, w0, w1;
};
__device__ __forceinline__ Taps make_taps(int Y, int X, int h, int w, float sh, float sw, bool ac) {
  int y0, y1, x0, x1;
  float ly, lx;
  resize_src(sh, Y, h, ac, y0, y1, ly);
  resize_src(sw, X, w, ac, x0, x1, lx);
  Taps t;
  t.o00 = y0 * w + x0; t.o01 = y0 * w + x1; t.o10 = y1 * w + x0; t.o11 = y1 * w + x1;
  t.h1 = ly; t.h0 = __fsub_rn(1.f, ly); t.w1 = lx; t.w0 = __fsub_rn(1.f, lx);
  t.w00 = t.h0 * t.w0; t.w01 = t.h0 * t.w1; t.w10 = t.h1 * t.w0; t.w11 = t.h1 * t.w1;
  return t;
}
// ATen's expression h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11) with its contraction pinned (common.cuh)
template <typename T> __device__ __forceinline__ float interp(const T* plane, const Taps& t) {
  float v00 = to_float<T>(plane[t.o00]), v01 = to_float<T>(plane[t.o01]);
  float v10 = to_float<T>(plane[t.o10]), v11 = to_float<T>(plane[t.o11]);
  return aten_bilerp(t.h0, t.h1, t.w0, t.w1, v00, v01, v10, v11);
}

template <typename T, int V, int CH, bool UP>
__global__ void __launch_bounds__(256, (V == 8 ? 2 : 3)) ce_fwd_kernel(const CeFwdParams p) {
  static_assert(!UP || V == 1, "resize-fused variant is one pixel per thread");
  const int n = blockIdx.y;
  const int C = p.C;
  const long long HW = (long long)p.H * p.W;
  const long long hw = (long long)p.h * p.w;
  const long long px0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;

  float loss_acc = 0.f;
  int n_valid = 0, n_correct = 0, n_bad = 0, n_acc = 0;
  // one-hot Dice terms of this thread's pixels (class clamped to [0,C-1], dice_loss.py:119-122)
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  const bool dice = p.dice_part != nullptr;
  int ycl[V];
  float pyv[V];
#pragma unroll
  for (int v = 0; v < V; ++v) { ycl[v] = -1; pyv[v] = 0.f; }
  if (dice) {
    float* bins = reinterpret_cast<float*>(smem_dyn);
    for (int i = threadIdx.x; i < 2 * 8 * C; i += blockDim.x) bins[i] = 0.f;
    __syncthreads();
  }

  if (px0 < HW) {
    long long y[V];
    load_labels<V>(p.labels, p.label_dtype, (size_t)n * HW + px0, y);

    float m[V], s[V];
    int idx[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { m[v] = neg_inf(); s[v] = 0.f; idx[v] = 0; }

    const T* img = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * hw;
    Taps tp;
    if constexpr (UP) tp = make_taps((int)(px0 / p.W), (int)(px0 % p.W), p.h, p.w, p.sh, p.sw, p.align_corners != 0);

    // chunk loader: CH classes x V pixels (streaming 128-bit loads; the resize-fused variant interpolates 4 taps)
    const T* cls_ptr = UP ? img : img + px0;   // running pointer over the class dimension (one 64-bit add per class)
    const long long cls_stride = UP ? hw : HW;
    // A chunk = CH classes x V pixels. Loads stay PACKED (RawVec) until they are reduced, so a bf16 thread keeps as
    // many bytes in flight per register as an fp32 one; the resize-fused variant interpolates its 4 taps instead.
    // Full chunks run predicate-free; the C % CH tail classes take the predicated flavour once.
    struct Chunk { RawVec<T, V> r[CH]; float f[UP ? CH : 1]; };
    auto load_chunk = [&](Chunk& ck, int left) {   // left >= CH for a full chunk
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (i < left) {
          if constexpr (UP) ck.f[i] = interp<T>(cls_ptr, tp);
          else ck.r[i] = load_raw<T, V>(cls_ptr);
        }
        cls_ptr += cls_stride;
      }
    };
    // online soft-max update with one chunk (running max m, rescaled sum s, arg-max idx)
    auto reduce_chunk = [&](int c0, const Chunk& ck, int left, bool full) {
      float z[CH][V];
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (full || i < left) {
          if constexpr (UP) z[i][0] = ck.f[i];
          else unpack_raw<T, V>(ck.r[i], z[i]);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) z[i][v] = neg_inf();
        }
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        // chunk max by a select-free tree, then its (lowest) index by equality selects: no predicate chains
        float cm = z[0][v];
#pragma unroll
        for (int i = 1; i < CH; ++i) cm = fmaxf(cm, z[i][v]);
        int li = CH - 1;
#pragma unroll
        for (int i = CH - 2; i >= 0; --i) li = (z[i][v] == cm) ? i : li;
        const bool up_max = cm > m[v];                       // strict '>': an earlier class keeps a tie
        idx[v] = up_max ? c0 + li : idx[v];
        const float nmx = up_max ? cm : m[v];
        const float nm = -nmx * kLog2e;
        float acc0 = s[v] * ex2(fmaf(m[v], kLog2e, nm)), acc1 = 0.f;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const float e = ex2(fmaf(z[i][v], kLog2e, nm));
          if (i & 1) acc1 += e;
          else acc0 += e;
        }
        s[v] = acc0 + acc1;
        m[v] = nmx;
      }
    };
    // double-buffered over the full chunks: the loads of chunk k+1 are in flight while chunk k is reduced
    {
      const int nfull = C / CH;
      Chunk za, zb;
      if (nfull > 0) load_chunk(za, CH);
      for (int k = 0; k < nfull; k += 2) {
        const bool has_b = k + 1 < nfull;
        if (has_b) load_chunk(zb, CH);
        reduce_chunk(k * CH, za, CH, true);
        if (has_b) {
          if (k + 2 < nfull) load_chunk(za, CH);
          reduce_chunk((k + 1) * CH, zb, CH, true);
        }
      }
      const int tail = C - nfull * CH;
      if (tail > 0) {
        load_chunk(za, tail);
        reduce_chunk(nfull * CH, za, tail, false);
      }
    }

    float lse[V], lpx[V], pwv[V];
    if (p.pw) {
      load_f32<V>(p.pw + (size_t)n * HW + px0, pwv);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) pwv[v] = 1.f;
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      lse[v] = m[v] + fast_log(s[v]);
      const long long yy = y[v];
      const bool ign = (yy == p.ignore_index);
      const bool inr = (yy >= 0 && yy < (long long)C);
      const bool valid = !ign && inr;
      n_bad += (!ign && !inr);
      n_valid += !ign;
      float l = 0.f;
      if (valid || dice) {
        const long long ycc = yy < 0 ? 0 : (yy >= (long long)C ? C - 1 : yy);
        float zy;
        if constexpr (UP) zy = interp<T>(img + (size_t)ycc * hw, tp);
        else zy = to_float<T>(img[(size_t)ycc * HW + px0 + v]);
        if (valid) {
          const float wt = p.cw ? __ldg(p.cw + yy) : 1.f;
          l = wt * (lse[v] - zy) * pwv[v];
        }
        if (dice) {
          const bool dv = (yy != p.dice_ignore);                                 // valid_mask
          ycl[v] = (p.tversky && !dv) ? -1 : (int)ycc;                           // Tversky counts sum t*v, Dice sum t
          pyv[v] = dv ? ex2((zy - lse[v]) * kLog2e) : 0.f;
        }
      }
      lpx[v] = l * p.lw;
      loss_acc += l;
      const bool av = p.acc_has_ignore ? (yy != p.acc_ignore) : true;
      n_acc += av;
      n_correct += (av && (long long)idx[v] == yy);
    }
    if (p.tversky) {   // every later pass forms p = exp(z - lse): +inf masks the ignored pixels out of all of them
#pragma unroll
      for (int v = 0; v < V; ++v) lse[v] = (y[v] == p.dice_ignore) ? __int_as_float(0x7f800000) : lse[v];
    }
    if (p.lse) store_f32<V>(p.lse + (size_t)n * HW + px0, lse);
    if (p.loss_px) store_f32<V>(p.loss_px + (size_t)n * HW + px0, lpx);
  }

  if (dice) {
    // warp-aggregated scatter of (p_y * valid, 1) into this warp's class bins, then one flush per CTA
    float* A_s = reinterpret_cast<float*>(smem_dyn);
    float* T_s = A_s + 8 * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      onehot_bins_add(A_s + warp * C, T_s + warp * C, ycl[v], pyv[v], lane);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a = 0.f, t = 0.f;
      for (int q = 0; q < 8; ++q) { a += A_s[q * C + c]; t += T_s[q * C + c]; }
      if (t != 0.f) {
        atomicAdd(p.dice_part + ((size_t)n * C + c) * 3 + 0, (double)a);
        atomicAdd(p.dice_part + ((size_t)n * C + c) * 3 + 2, (double)t);
      }
    }
  }
  cta_flush_stats(loss_acc, n_valid, n_correct, n_bad, n_acc, p.stats);
}

// ------------------------------------------------------------------------------------------------
struct CeBwdParams {
  const void* logits;
  const void* labels;
  const float* pw;
  const float* cw;
  const float* lse;
  const float* grad_out;   // scalar or null
  const float* grad_px;    // per pixel or null
  const unsigned long long* stats;
  void* grad;
  int label_dtype;
  int N, C, h, w, H, W;
  int align_corners;
  int use_nvalid;
  long long ignore_index;
  float scale_host;
  float sh, sw;
};

__device__ __forceinline__ float ce_global_scale(const CeBwdParams& p) {
  float G = p.scale_host;
  if (p.grad_out) G *= __ldg(p.grad_out);
  if (p.use_nvalid) {
    const double nv = (double)(long long)p.stats[B200SEG_ST_N_VALID];
    G = (float)((double)G / (nv + 1.1920928955078125e-07));
  }
  return G;
}

// Label-resolution logits only. A resize-fused backward lives in loss_upgen.cuh (deterministic cell-owner sums); the
// atomicAdd scatter this kernel once carried for other ratios (ATen's own non-deterministic design) is gone: shapes the
// cell-owner kernel does not take are resized first (csrc/resize.cu, deterministic gather backward).
template <typename T, int V, int CH>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const CeBwdParams p) {
  // reverse launch order: the forward kernel read the batch front to back and its tail is still in the 126 MB L2
  const int n = gridDim.y - 1 - blockIdx.y;
  const int C = p.C;
  const long long HW = (long long)p.H * p.W;
  const long long hw = (long long)p.h * p.w;
  const long long px0 = ((long long)(gridDim.x - 1 - blockIdx.x) * blockDim.x + threadIdx.x) * V;
  if (px0 >= HW) return;
  const float G = ce_global_scale(p);

  long long y[V];
  load_labels<V>(p.labels, p.label_dtype, (size_t)n * HW + px0, y);
  float lse[V], coef[V], nl[V];
  load_f32<V>(p.lse + (size_t)n * HW + px0, lse);
  if (p.pw) {
    load_f32<V>(p.pw + (size_t)n * HW + px0, coef);
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v) coef[v] = 1.f;
  }
  if (p.grad_px) {
    float g[V];
    load_f32<V>(p.grad_px + (size_t)n * HW + px0, g);
#pragma unroll
    for (int v = 0; v < V; ++v) coef[v] *= g[v];
  }
  int yc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const long long yy = y[v];
    const bool valid = (yy != p.ignore_index) && yy >= 0 && yy < (long long)C;
    const float wt = (valid && p.cw) ? __ldg(p.cw + yy) : 1.f;
    coef[v] = valid ? coef[v] * wt * G : 0.f;
    yc[v] = valid ? (int)yy : -1;
    nl[v] = -lse[v] * kLog2e;
  }

  const T* img = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * hw;
  T* gq = reinterpret_cast<T*>(p.grad) + (size_t)n * C * HW + px0;
  const T* q = img + px0;
  for (int c0 = 0; c0 < C; c0 += CH) {
    RawVec<T, V> raw[CH];
    const int left = C - c0;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (i < left) raw[i] = load_raw<T, V>(q);
      q += HW;
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (i < left) {
        float g[V], zz[V];
        unpack_raw<T, V>(raw[i], zz);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          g[v] = coef[v] * ex2(fmaf(zz[v], kLog2e, nl[v]));
          if (c0 + i == yc[v]) g[v] -= coef[v];
        }
        store_vec<T, V>(gq, g);
      }
      gq += HW;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// finalize: statistics -> scalars (models/losses/utils.py:48-80, accuracy.py:51-60, dice_loss.py:31-58)
__global__ void __launch_bounds__(1024) finalize_kernel(const b200seg_finalize_desc d) {
  __shared__ double sred[32];
  pdl_wait();      // scheduled under the tail of the loss kernel (programmatic dependent launch)
  const double eps = 1.1920928955078125e-07;  // torch.finfo(torch.float32).eps
  if (threadIdx.x == 0) {
    const double sum = *reinterpret_cast<const double*>(d.stats + B200SEG_ST_CE_SUM);
    const double n_valid = (double)(long long)d.stats[B200SEG_ST_N_VALID];
    const double n_correct = (double)(long long)d.stats[B200SEG_ST_N_CORRECT];
    const double n_acc = (double)(long long)d.stats[B200SEG_ST_N_ACC];
    double loss = sum;
    if (d.ce_reduction == B200SEG_RED_MEAN) {
      if (d.ce_has_avg_factor) loss = sum / (double)(float)(d.ce_avg_factor + eps);
      else if (d.ce_avg_non_ignore) loss = sum / (double)(float)(n_valid + eps);
      else loss = sum / (double)d.n_pixels;
    }
    *d.out_loss_ce = (float)((double)d.ce_loss_weight * loss);
    // accuracy.py:55-60: (correct.float().sum() + eps) * (100.0 / (n + eps)), evaluated in fp32
    const float r = (float)(100.0 / (n_acc + eps));
    if (d.out_acc) *d.out_acc = ((float)n_correct + (float)eps) * r;
    if (d.log_vec) {
      d.log_vec[B200SEG_LOG_CE_SUM] = sum;
      d.log_vec[B200SEG_LOG_N_VALID] = n_valid;
      d.log_vec[B200SEG_LOG_N_CORRECT] = n_correct;
      d.log_vec[B200SEG_LOG_N_ACC] = n_acc;
      d.log_vec[B200SEG_LOG_N_BAD] = (double)(long long)d.stats[B200SEG_ST_N_BAD];
      d.log_vec[B200SEG_LOG_N_PIXELS] = (double)d.n_pixels;
      d.log_vec[B200SEG_LOG_DICE_SUM] = 0.0;
      d.log_vec[B200SEG_LOG_N_IMAGES] = (double)d.N;
    }
  }
  if (d.dice_part == nullptr) {
    if (threadIdx.x == 0 && d.out_loss_dice) *d.out_loss_dice = 0.f;
    return;
  }
  // K: d(loss_dice)/d(per-sample, per-class dice term)
  double K = (double)d.dice_loss_weight / ((double)d.C * (double)d.N);
  if (d.dice_has_avg_factor && d.dice_reduction == B200SEG_RED_MEAN) K /= (double)(float)(d.dice_avg_factor + eps);
  double part = 0.0;
  for (int i = threadIdx.x; i < d.N * d.C; i += blockDim.x) {
    const int c = i % d.C;
    const double A = d.dice_part[(size_t)i * 3 + 0], B = d.dice_part[(size_t)i * 3 + 1], T = d.dice_part[(size_t)i * 3 + 2];
    const double cwv = d.dice_class_weight ? (double)d.dice_class_weight[c] : 1.0;
    const bool skip = ((long long)c == d.dice_ignore_index);
    if (d.dice_mode == B200SEG_MODE_TVERSKY) {
      // tversky_loss.py:52-68: TP = A, FP = B - A, FN = T - A; 1 - (TP + s) / (TP + a FP + b FN + s)
      const double ta = (double)d.tversky_alpha, tb = (double)d.tversky_beta, sm = (double)d.dice_smooth;
      const double num = A + sm;
      const double den = (1.0 - ta - tb) * A + ta * B + tb * T + sm;
      if (!skip) part += cwv * (1.0 - num / den);
      if (d.dice_coef) {   // alpha = -dL/dA, beta = dL/dB (exponent 1)
        d.dice_coef[(size_t)i * 2 + 0] = skip ? 0.f : (float)(K * cwv * (1.0 / den - num * (1.0 - ta - tb) / (den * den)));
        d.dice_coef[(size_t)i * 2 + 1] = skip ? 0.f : (float)(K * cwv * num * ta / (den * den));
      }
      continue;
    }
    const double num = 2.0 * A + (double)d.dice_smooth;
    const double den = B + T + (double)d.dice_smooth;
    if (!skip) part += cwv * (1.0 - num / den);
    if (d.dice_coef) {
      d.dice_coef[(size_t)i * 2 + 0] = skip ? 0.f : (float)(K * cwv * 2.0 / den);
      d.dice_coef[(size_t)i * 2 + 1] = skip ? 0.f : (float)(K * cwv * num / (den * den));
    }
  }
  double r[1] = {part};
  block_sum<double, 1>(r, sred);
  if (threadIdx.x == 0) {
    if (d.out_loss_dice) *d.out_loss_dice = (float)(K * r[0]);
    // sum over (n,c) of cw_c * (1 - num/den): additive over images, so ranks can all-reduce it
    if (d.log_vec) d.log_vec[B200SEG_LOG_DICE_SUM] = r[0];
  }
}

// ------------------------------------------------------------------------------------------------
// host dispatch
template <typename T> static int launch_ce_fwd(const CeFwdParams& p, bool up, bool vec, cudaStream_t st) {
  const long long HW = (long long)p.H * p.W;
  constexpr int VV = 4;   // 4 pixels per thread for every dtype: 16-byte (fp32) / 8-byte (16-bit) loads; 8 pixels of
                          // per-thread soft-max state cost the 16-bit variant its occupancy (128 regs, 31 % of roofline)
  const size_t sm = p.dice_part ? (size_t)2 * 8 * p.C * sizeof(float) : 0;   // <= 32 KB for C <= 512
  if (up) {
    dim3 grid((unsigned)((HW + 255) / 256), p.N);
    ce_fwd_kernel<T, 1, 4, true><<<grid, 256, sm, st>>>(p);
  } else if (vec) {
    dim3 grid((unsigned)((HW / VV + 255) / 256), p.N);
    ce_fwd_kernel<T, VV, (sizeof(T) == 2 ? 8 : 4), false><<<grid, 256, sm, st>>>(p);
  } else {
    dim3 grid((unsigned)((HW + 255) / 256), p.N);
    ce_fwd_kernel<T, 1, 8, false><<<grid, 256, sm, st>>>(p);
  }
  count_launch();
  return check_launch("ce_fwd_kernel");
}

template <typename T> static int launch_ce_bwd(const CeBwdParams& p, bool vec, cudaStream_t st) {
  const long long HW = (long long)p.H * p.W;
  constexpr int VV = 16 / (int)sizeof(T);
  if (vec) {
    dim3 grid((unsigned)((HW / VV + 255) / 256), p.N);
    ce_bwd_kernel<T, VV, 8><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid((unsigned)((HW + 255) / 256), p.N);
    ce_bwd_kernel<T, 1, 8><<<grid, 256, 0, st>>>(p);
  }
  count_launch();
  return check_launch("ce_bwd_kernel");
}

int ce_fwd_dispatch(const b200seg_loss_desc* d, cudaStream_t st) {
  CeFwdParams p;
  p.logits = d->logits; p.labels = d->labels; p.pw = d->pixel_weight; p.cw = d->ce_class_weight;
  p.lse = (d->flags & B200SEG_WANT_LSE) ? d->lse : nullptr;
  p.loss_px = (d->flags & B200SEG_WANT_LOSS_PX) ? d->loss_px : nullptr;
  p.stats = reinterpret_cast<unsigned long long*>(d->stats);
  p.label_dtype = d->label_dtype;
  p.N = d->N; p.C = d->C; p.h = d->h; p.w = d->w; p.H = d->H; p.W = d->W;
  p.align_corners = d->align_corners; p.flags = d->flags;
  p.ignore_index = d->ignore_index; p.acc_has_ignore = d->acc_has_ignore; p.acc_ignore = d->acc_ignore_index;
  p.lw = d->ce_loss_weight;
  p.sh = resize_scale(d->h, d->H, d->align_corners != 0);
  p.sw = resize_scale(d->w, d->W, d->align_corners != 0);
  p.dice_part = (d->flags & B200SEG_WANT_DICE) ? d->dice_part : nullptr;
  p.dice_ignore = d->dice_ignore_index;
  p.tversky = (d->flags & B200SEG_WANT_DICE) && d->dice_mode == B200SEG_MODE_TVERSKY;
  const bool up = (d->h != d->H) || (d->w != d->W);
  const long long HW = (long long)d->H * d->W;
  const int VV = 4;
  const bool vec = !up && (HW % VV == 0) && aligned16(d->logits) && aligned16(d->labels) &&
                   (!p.pw || aligned16(p.pw)) && (!p.lse || aligned16(p.lse)) && (!p.loss_px || aligned16(p.loss_px));
  switch (d->logit_dtype) {
    case B200SEG_F32: return launch_ce_fwd<float>(p, up, vec, st);
    case B200SEG_BF16: return launch_ce_fwd<__nv_bfloat16>(p, up, vec, st);
    case B200SEG_F16: return launch_ce_fwd<__half>(p, up, vec, st);
  }
  set_error("unsupported logit dtype %d", d->logit_dtype);
  return 1;
}

int ce_bwd_dispatch(const b200seg_loss_bwd_desc* d, cudaStream_t st) {
  CeBwdParams p;
  p.logits = d->logits; p.labels = d->labels; p.pw = d->pixel_weight; p.cw = d->ce_class_weight;
  p.lse = d->lse; p.grad_out = d->ce_grad_out; p.grad_px = d->ce_grad_px;
  p.stats = reinterpret_cast<const unsigned long long*>(d->stats);
  p.grad = d->grad_logits;
  p.label_dtype = d->label_dtype;
  p.N = d->N; p.C = d->C; p.h = d->h; p.w = d->w; p.H = d->H; p.W = d->W;
  p.align_corners = d->align_corners; p.use_nvalid = d->ce_use_nvalid;
  p.ignore_index = d->ignore_index; p.scale_host = d->ce_scale_host;
  p.sh = resize_scale(d->h, d->H, d->align_corners != 0);
  p.sw = resize_scale(d->w, d->W, d->align_corners != 0);
  const bool up = (d->h != d->H) || (d->w != d->W);
  B200SEG_REQUIRE(!up, "loss_bwd: logits must be at label resolution — the resize-fused backward is b200seg_loss_fused_fwdbwd "
                       "(H >= h, W >= w, C <= 32); resize other shapes first (b200seg_resize_bilinear_fwd / _bwd)");
  const long long HW = (long long)d->H * d->W;
  const int VV = 16 / logit_bytes(d->logit_dtype);
  const bool vec = (HW % VV == 0) && aligned16(d->logits) && aligned16(d->labels) && aligned16(d->lse) &&
                   aligned16(d->grad_logits) && (!p.pw || aligned16(p.pw)) && (!p.grad_px || aligned16(p.grad_px));
  switch (d->logit_dtype) {
    case B200SEG_F32: return launch_ce_bwd<float>(p, vec, st);
    case B200SEG_BF16: return launch_ce_bwd<__nv_bfloat16>(p, vec, st);
    case B200SEG_F16: return launch_ce_bwd<__half>(p, vec, st);
  }
  set_error("unsupported logit dtype %d", d->logit_dtype);
  return 1;
}

int finalize_dispatch(const b200seg_finalize_desc* d, cudaStream_t st) {
  // one CTA; 1024 threads when there are thousands of (sample, class) Dice terms to walk (fp64 divisions), else 256
  const int threads = (d->dice_part != nullptr && (long long)d->N * d->C > 512) ? 1024 : 256;
  launch_pdl(finalize_kernel, dim3(1), dim3(threads), 0, st, *d);
  count_launch();
  return check_launch("finalize_kernel");
}

}  // namespace b200seg
