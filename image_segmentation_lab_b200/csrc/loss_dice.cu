// Dice for MANY classes (C > 32) as pure streaming kernels, sm_100a.
//
// Replaces DiceLoss.forward models/losses/dice_loss.py:103-134 (F.softmax, the int64 (N,H,W,C) one-hot — 5 GB at
// ADE20K shape — and the 150-iteration dice_loss / binary_dice_loss loop :23-58) and its autograd backward.
//
// Holding 150 soft-max values per pixel on chip needs either a class split with inter-warp exchanges (barrier
// coupled, ~15 warps/SM: measured 9 % of the HBM roofline) or shared-memory staging that exceeds the MUFU / shared
// bandwidth budget at bf16. Instead the work is cut into streams that each run near the HBM roofline:
//   forward  A  ce_fwd_kernel (loss_stream.cu): online soft-max -> lse per pixel, CE, accuracy, one-hot Dice sums
//            B  dice_sumsq_kernel: CLASS-major. A warp owns CW classes, a lane owns V pixels; p = 2^((z-lse)log2e)
//               needs no cross-thread data, and sum_px p^e accumulates in one register per owned class across all
//               the tiles the CTA visits (one shuffle tree per class per CTA lifetime, no barriers, no smem).
//   backward C  dice_dot_kernel:  pixel-major, dot_px = sum_c p_c g_c (g = dL/dp) -> (N,H,W) scratch
//            D  dice_grad_kernel: pixel-major, grad = p (g - dot) + k (p - onehot), streamed store
// Traffic: forward 2 reads, backward 2 reads + 1 write of the logits (algorithmic: 1 and 1 + 1), one MUFU.EX2 per
// element per pass. Roofline: HBM.
#include <cstdlib>

#include "common.cuh"

namespace b200seg {

struct DiceParams {
  const void* logits;
  const void* labels;
  const float* pw;
  const float* cw;
  const float* lse;
  const float* dice_coef;     // (N,C,2) [alpha, beta]
  const float* dice_grad_out;
  const float* ce_grad_out;
  const float* ce_grad_px;
  const unsigned long long* stats;
  double* dice_part;
  float* dot;                 // (N,H,W) scratch
  void* grad;
  int label_dtype;
  int N, C;
  long long HW;
  int flags;
  long long ignore_index;
  long long dice_ignore;
  float dice_exponent;
  float ce_scale_host;
  int ce_use_nvalid;
  int tiles;
};

// ------------------------------------------------------------------------------------------------ B: sum_px p_c^e
template <typename T, int V, int CW, int CH>
__global__ void __launch_bounds__(256) dice_sumsq_kernel(const DiceParams p) {
  static_assert(CW % CH == 0 && CW <= 32, "class window");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // images and tiles are walked in REVERSE launch order: this kernel follows ce_fwd_kernel, which leaves the last
  // ~100 MB it read (the tail of the batch) in the 126 MB L2
  const int n = gridDim.y - 1 - blockIdx.y;
  const int C = p.C;
  const long long HW = p.HW;
  const int c0 = (blockIdx.z * 8 + warp) * CW;          // first class of this warp
  const int ncls_w = C - c0 < CW ? C - c0 : CW;         // may be <= 0
  const bool e2 = (p.dice_exponent == 2.f), e1 = (p.dice_exponent == 1.f);
  const T* img = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * HW;

  float acc[CW];
#pragma unroll
  for (int i = 0; i < CW; ++i) acc[i] = 0.f;

  if (ncls_w > 0) {
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      const int tile = p.tiles - 1 - t;
      const long long px0 = ((long long)tile * 32 + lane) * V;
      if (px0 >= HW) continue;
      float nl[V];
      {
        float lse[V];
        load_px_f32<V>(p.lse + (size_t)n * HW + px0, lse);
#pragma unroll
        for (int v = 0; v < V; ++v) nl[v] = -lse[v] * kLog2e;
      }
      const T* q = img + (size_t)c0 * HW + px0;
      struct Chunk { RawVec<T, V> r[CH]; };
      Chunk za, zb;
      auto load_chunk = [&](int k0, Chunk& ck) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          if (k0 + i < ncls_w) ck.r[i] = load_raw<T, V>(q);
          q += HW;
        }
      };
      auto consume = [&](int k0, const Chunk& ck, float (&a)[CH]) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          a[i] = 0.f;
          if (k0 + i < ncls_w) {
            float z[V];
            unpack_raw<T, V>(ck.r[i], z);
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float pr = ex2(fmaf(z[v], kLog2e, nl[v]));
              a[i] += e2 ? pr * pr : (e1 ? pr : (pr > 0.f ? __powf(pr, p.dice_exponent) : 0.f));
            }
          }
        }
      };
      load_chunk(0, za);
#pragma unroll
      for (int k0 = 0; k0 < CW; k0 += 2 * CH) {
        float a[CH];
        if (k0 + CH < CW) load_chunk(k0 + CH, zb);
        consume(k0, za, a);
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[k0 + i] += a[i];
        if (k0 + CH < CW) {
          if (k0 + 2 * CH < CW) load_chunk(k0 + 2 * CH, za);
          consume(k0 + CH, zb, a);
#pragma unroll
          for (int i = 0; i < CH; ++i) acc[k0 + CH + i] += a[i];
        }
      }
    }
  }
  float mine = 0.f;
#pragma unroll
  for (int i = 0; i < CW; ++i) {
    const float tot = warp_sum(acc[i]);
    if (lane == i) mine = tot;
  }
  if (lane < ncls_w) atomicAdd(p.dice_part + ((size_t)n * C + c0 + lane) * 3 + 1, (double)mine);
}

// ------------------------------------------------------------------------------------------------ C: per-pixel dot
// dot_px = sum_c p_c * g_c with g_c = e*beta_c*p_c^(e-1) - alpha_c * onehot_c * valid   (upstream gradient folded in)
template <typename T, int V, int CH>
__global__ void __launch_bounds__(256) dice_dot_kernel(const DiceParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* beta_s = reinterpret_cast<float*>(smem_raw);   // [C] e * god * beta
  const int n = blockIdx.y;
  const int C = p.C;
  const long long HW = p.HW;
  const float god = p.dice_grad_out ? __ldg(p.dice_grad_out) : 1.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    beta_s[c] = p.dice_exponent * god * __ldg(p.dice_coef + ((size_t)n * C + c) * 2 + 1);
  __syncthreads();
  const long long px0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (px0 >= HW) return;
  const bool e2 = (p.dice_exponent == 2.f), e1 = (p.dice_exponent == 1.f);
  const T* img = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * HW;
  float nl[V], nl2[V], dot[V];
  {
    float lse[V];
    if constexpr (V == 8) {
      float a[4], b[4];
      load_vec<float, 4>(p.lse + (size_t)n * HW + px0, a);
      load_vec<float, 4>(p.lse + (size_t)n * HW + px0 + 4, b);
#pragma unroll
      for (int k = 0; k < 4; ++k) { lse[k] = a[k]; lse[4 + k] = b[k]; }
    } else {
      load_vec<float, V>(p.lse + (size_t)n * HW + px0, lse);
    }
#pragma unroll
    for (int v = 0; v < V; ++v) { nl[v] = -lse[v] * kLog2e; nl2[v] = 2.f * nl[v]; dot[v] = 0.f; }
  }
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
        const float b = beta_s[c0 + i];
        float zz[V];
        unpack_raw<T, V>(raw[i], zz);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          if (e2) {   // p^2 = 2^(2 (z - lse) log2e): one EX2, no multiply
            dot[v] = fmaf(b, ex2(fmaf(zz[v], 2.f * kLog2e, nl2[v])), dot[v]);
          } else {
            const float pr = ex2(fmaf(zz[v], kLog2e, nl[v]));
            if (e1) dot[v] = fmaf(b, pr, dot[v]);
            else dot[v] += pr > 0.f ? b * __powf(pr, p.dice_exponent) : 0.f;
          }
        }
      }
    }
  }
  long long y[V];   // only needed for the one-hot term: loaded after the class loop to keep 2*V registers free in it
  load_labels<V>(p.labels, p.label_dtype, (size_t)n * HW + px0, y);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const long long yy = y[v];
    if (yy != p.dice_ignore) {
      const long long ycc = yy < 0 ? 0 : (yy >= (long long)C ? C - 1 : yy);
      const float a = god * __ldg(p.dice_coef + ((size_t)n * C + ycc) * 2 + 0);
      const float zy = to_float<T>(img[(size_t)ycc * HW + px0 + v]);
      dot[v] -= a * ex2(fmaf(zy, kLog2e, nl[v]));
    }
  }
  if constexpr (V == 8) {
    float a[4] = {dot[0], dot[1], dot[2], dot[3]}, b[4] = {dot[4], dot[5], dot[6], dot[7]};
    store_vec<float, 4>(p.dot + (size_t)n * HW + px0, a);
    store_vec<float, 4>(p.dot + (size_t)n * HW + px0 + 4, b);
  } else {
    store_vec<float, V>(p.dot + (size_t)n * HW + px0, dot);
  }
}

// ------------------------------------------------------------------------------------------------ D: gradient
template <typename T, int V, int CH>
__global__ void __launch_bounds__(256) dice_grad_kernel(const DiceParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* beta_s = reinterpret_cast<float*>(smem_raw);
  // reverse launch order: dice_dot_kernel has just read the batch front to back, its tail is still in L2
  const int n = gridDim.y - 1 - blockIdx.y;
  const int C = p.C;
  const long long HW = p.HW;
  const float god = p.dice_grad_out ? __ldg(p.dice_grad_out) : 1.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    beta_s[c] = p.dice_exponent * god * __ldg(p.dice_coef + ((size_t)n * C + c) * 2 + 1);
  __syncthreads();
  const long long px0 = ((long long)(gridDim.x - 1 - blockIdx.x) * blockDim.x + threadIdx.x) * V;
  if (px0 >= HW) return;
  const bool e2 = (p.dice_exponent == 2.f), e1 = (p.dice_exponent == 1.f);
  const bool want_ce = (p.flags & B200SEG_WANT_CE) != 0;
  float Gce = 0.f;
  if (want_ce) {
    Gce = p.ce_scale_host;
    if (p.ce_grad_out) Gce *= __ldg(p.ce_grad_out);
    if (p.ce_use_nvalid) {
      const double nv = (double)(long long)p.stats[B200SEG_ST_N_VALID];
      Gce = (float)((double)Gce / (nv + 1.1920928955078125e-07));
    }
  }
  const T* img = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * HW;
  auto load_px = [&](const float* base, float (&o)[V]) {
    if constexpr (V == 8) {
      float a[4], b[4];
      load_vec<float, 4>(base, a);
      load_vec<float, 4>(base + 4, b);
#pragma unroll
      for (int k = 0; k < 4; ++k) { o[k] = a[k]; o[4 + k] = b[k]; }
    } else {
      load_vec<float, V>(base, o);
    }
  };
  float nl[V], sub[V], kk[V], da[V];
  int ycl[V];
  {
    float lse[V], dot[V], pwv[V], gpx[V];
    load_px(p.lse + (size_t)n * HW + px0, lse);
    load_px(p.dot + (size_t)n * HW + px0, dot);
#pragma unroll
    for (int v = 0; v < V; ++v) { pwv[v] = 1.f; gpx[v] = 1.f; }
    if (want_ce && p.pw) load_px(p.pw + (size_t)n * HW + px0, pwv);
    if (want_ce && p.ce_grad_px) load_px(p.ce_grad_px + (size_t)n * HW + px0, gpx);
    long long y[V];
    load_labels<V>(p.labels, p.label_dtype, (size_t)n * HW + px0, y);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      nl[v] = -lse[v] * kLog2e;
      const long long yy = y[v];
      const long long ycc = yy < 0 ? 0 : (yy >= (long long)C ? C - 1 : yy);
      ycl[v] = (int)ycc;
      da[v] = (yy != p.dice_ignore) ? god * __ldg(p.dice_coef + ((size_t)n * C + ycc) * 2 + 0) : 0.f;
      const bool valid = (yy != p.ignore_index) && yy >= 0 && yy < (long long)C;
      kk[v] = (want_ce && valid) ? Gce * pwv[v] * gpx[v] * (p.cw ? __ldg(p.cw + yy) : 1.f) : 0.f;
      sub[v] = kk[v] - dot[v];
    }
  }
  const T* q = img + px0;
  T* gq = reinterpret_cast<T*>(p.grad) + (size_t)n * C * HW + px0;
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
        const float b = beta_s[c0 + i];
        float g[V], zz[V];
        unpack_raw<T, V>(raw[i], zz);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float pr = ex2(fmaf(zz[v], kLog2e, nl[v]));
          float gd;
          if (e2) gd = b * pr;
          else if (e1) gd = b;
          else gd = pr > 0.f ? b * __powf(pr, p.dice_exponent - 1.f) : 0.f;
          g[v] = pr * (gd + sub[v]);
        }
        store_vec<T, V>(gq, g);
      }
      gq += HW;
    }
  }
  // one-hot terms: instead of a compare per element, the label's class is re-stored by the same thread (program order)
  // with its extra -(p_y alpha_y + k); p_y is recomputed from the label's logit with the operations of the class loop
  T* gimg = reinterpret_cast<T*>(p.grad) + (size_t)n * C * HW;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    if (da[v] != 0.f || kk[v] != 0.f) {
      const int yc = ycl[v];
      const float pr = ex2(fmaf(to_float<T>(img[(size_t)yc * HW + px0 + v]), kLog2e, nl[v]));
      const float b = beta_s[yc];
      float gd;
      if (e2) gd = b * pr;
      else if (e1) gd = b;
      else gd = pr > 0.f ? b * __powf(pr, p.dice_exponent - 1.f) : 0.f;
      gimg[(size_t)yc * HW + px0 + v] = from_float<T>(pr * (gd + sub[v]) - fmaf(pr, da[v], kk[v]));
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
template <typename T> static int dice_fwd_t(const DiceParams& p0, bool vec, cudaStream_t st) {
  DiceParams p = p0;
  constexpr int VB = 16 / (int)sizeof(T);  // 16-byte loads kept packed: 512 B per warp per class row
  constexpr int CW = 20;
  const int zb = (p.C + 8 * CW - 1) / (8 * CW);
  if (vec) {
    p.tiles = (int)((p.HW + 32 * VB - 1) / (32 * VB));
    int gx = (kSMs * 4 + p.N * zb - 1) / (p.N * zb);
    if (gx > p.tiles) gx = p.tiles;
    dim3 grid(gx < 1 ? 1 : gx, p.N, zb);
    dice_sumsq_kernel<T, VB, CW, 4><<<grid, 256, 0, st>>>(p);
  } else {
    p.tiles = (int)((p.HW + 31) / 32);
    int gx = (kSMs * 4 + p.N * zb - 1) / (p.N * zb);
    if (gx > p.tiles) gx = p.tiles;
    dim3 grid(gx < 1 ? 1 : gx, p.N, zb);
    dice_sumsq_kernel<T, 1, CW, 4><<<grid, 256, 0, st>>>(p);
  }
  count_launch();
  return check_launch("dice_sumsq_kernel");
}

template <typename T> static int dice_bwd_t(const DiceParams& p, bool vec, cudaStream_t st) {
  constexpr int VV = 16 / (int)sizeof(T);
  const size_t sm = (size_t)p.C * sizeof(float);
  if (vec) {
    dim3 grid((unsigned)((p.HW / VV + 255) / 256), p.N);
    dice_dot_kernel<T, VV, 8><<<grid, 256, sm, st>>>(p);
    count_launch();
    if (int e = check_launch("dice_dot_kernel")) return e;
    dice_grad_kernel<T, VV, 8><<<grid, 256, sm, st>>>(p);
  } else {
    dim3 grid((unsigned)((p.HW + 255) / 256), p.N);
    dice_dot_kernel<T, 1, 8><<<grid, 256, sm, st>>>(p);
    count_launch();
    if (int e = check_launch("dice_dot_kernel")) return e;
    dice_grad_kernel<T, 1, 8><<<grid, 256, sm, st>>>(p);
  }
  count_launch();
  return check_launch("dice_grad_kernel");
}

int ce_fwd_dispatch(const b200seg_loss_desc* d, cudaStream_t st);
int cs_fwd_dispatch(const b200seg_loss_desc* d, cudaStream_t st);
int cs_bwd_dispatch(const b200seg_loss_bwd_desc* d, cudaStream_t st);
bool cs_supported(const void* logits, const void* labels, const void* lse, const void* grad, int logit_dtype, int label_dtype,
                  int C, long long HW, float dice_exponent, int dice_mode, bool want_dice);
// B200SEG_NO_CS=1 keeps the five-pass streaming kernels (A/B measurements and tests of the fallback)
static bool cs_disabled() {
  const char* e = getenv("B200SEG_NO_CS");
  return e && e[0] == '1';
}

// forward for C > 32: A (CE stream kernel with the one-hot Dice sums) then B
int dice_stream_fwd_dispatch(const b200seg_loss_desc* d, cudaStream_t st) {
  B200SEG_REQUIRE(d->h == d->H && d->w == d->W, "dice forward needs logits at label resolution (resize first)");
  B200SEG_REQUIRE(d->C <= 512, "dice path supports at most 512 classes (got %d)", d->C);
  B200SEG_REQUIRE(d->dice_part != nullptr, "dice_part workspace is NULL");
  B200SEG_REQUIRE(d->dice_exponent > 0.f, "dice exponent must be > 0");
  B200SEG_REQUIRE((d->flags & B200SEG_WANT_LSE) && d->lse, "dice forward with more than 32 classes needs the lse buffer");
  // one read of the logits (class-sliced bulk-copy pipeline, loss_cs.cu) when the tensors are 16-byte tileable
  if (!cs_disabled() && !(d->flags & B200SEG_WANT_LOSS_PX) &&
      cs_supported(d->logits, d->labels, d->lse, nullptr, d->logit_dtype, d->label_dtype, d->C, (long long)d->H * d->W,
                   d->dice_exponent, d->dice_mode, true))
    return cs_fwd_dispatch(d, st);
  if (int e = ce_fwd_dispatch(d, st)) return e;
  DiceParams p = {};
  p.logits = d->logits; p.lse = d->lse; p.dice_part = d->dice_part;
  p.N = d->N; p.C = d->C; p.HW = (long long)d->H * d->W;
  p.dice_exponent = d->dice_exponent;
  const int VB = 16 / logit_bytes(d->logit_dtype);
  const bool vec = (p.HW % VB == 0) && aligned16(d->logits) && aligned16(d->lse);
  switch (d->logit_dtype) {
    case B200SEG_F32: return dice_fwd_t<float>(p, vec, st);
    case B200SEG_BF16: return dice_fwd_t<__nv_bfloat16>(p, vec, st);
    case B200SEG_F16: return dice_fwd_t<__half>(p, vec, st);
  }
  set_error("unsupported logit dtype %d", d->logit_dtype);
  return 1;
}

int dice_stream_bwd_dispatch(const b200seg_loss_bwd_desc* d, cudaStream_t st) {
  B200SEG_REQUIRE(d->h == d->H && d->w == d->W, "dice backward needs logits at label resolution");
  B200SEG_REQUIRE(d->dice_coef != nullptr && d->lse != nullptr, "dice backward needs dice_coef and lse");
  if (!cs_disabled() && !d->ce_grad_px && (d->flags & B200SEG_WANT_DICE) &&
      cs_supported(d->logits, d->labels, d->lse, d->grad_logits, d->logit_dtype, d->label_dtype, d->C, (long long)d->H * d->W,
                   d->dice_exponent, d->dice_mode, true) &&
      (!d->pixel_weight || aligned16(d->pixel_weight)))
    return cs_bwd_dispatch(d, st);
  B200SEG_REQUIRE(d->scratch_px != nullptr, "dice backward with more than 32 classes needs the (N,H,W) f32 scratch_px");
  DiceParams p = {};
  p.logits = d->logits; p.labels = d->labels; p.pw = d->pixel_weight; p.cw = d->ce_class_weight;
  p.lse = d->lse; p.dice_coef = d->dice_coef; p.dice_grad_out = d->dice_grad_out;
  p.ce_grad_out = d->ce_grad_out; p.ce_grad_px = d->ce_grad_px;
  p.stats = reinterpret_cast<const unsigned long long*>(d->stats);
  p.dot = d->scratch_px; p.grad = d->grad_logits;
  p.label_dtype = d->label_dtype;
  p.N = d->N; p.C = d->C; p.HW = (long long)d->H * d->W;
  p.flags = d->flags; p.ignore_index = d->ignore_index; p.dice_ignore = d->dice_ignore_index;
  p.dice_exponent = d->dice_exponent; p.ce_scale_host = d->ce_scale_host; p.ce_use_nvalid = d->ce_use_nvalid;
  const int VV = 16 / logit_bytes(d->logit_dtype);
  const bool vec = (p.HW % VV == 0) && aligned16(d->logits) && aligned16(d->labels) && aligned16(d->lse) &&
                   aligned16(d->grad_logits) && aligned16(d->scratch_px) && (!p.pw || aligned16(p.pw)) &&
                   (!p.ce_grad_px || aligned16(p.ce_grad_px));
  switch (d->logit_dtype) {
    case B200SEG_F32: return dice_bwd_t<float>(p, vec, st);
    case B200SEG_BF16: return dice_bwd_t<__nv_bfloat16>(p, vec, st);
    case B200SEG_F16: return dice_bwd_t<__half>(p, vec, st);
  }
  set_error("unsupported logit dtype %d", d->logit_dtype);
  return 1;
}

}  // namespace b200seg
