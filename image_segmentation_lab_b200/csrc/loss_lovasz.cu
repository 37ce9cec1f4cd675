// Lovasz-Softmax / Lovasz hinge loss (the tail of SURVEY 8 f4), sm_100a.
//
// Replaces models/losses/lovasz_loss.py:26-234. Per class c the reference materialises softmax(N,C,H,W), permutes it to
// (P,C), compacts the valid pixels (boolean index + nonzero), then for every class builds fg = (labels == c),
// errors = |fg - p_c|, torch.sort(errors, descending), gathers fg through the permutation, forms the Jaccard gradient
// with two fp32 cumsums and a shifted difference (lovasz_grad, :26-39) and takes a dot product: ~12 launches and ~10
// (P,)-sized temporaries per class in a Python loop, a host sync per class for 'present' (:153), and an autograd graph
// that walks all of it backwards.
//
// Here, per (image group g, class c) SEGMENT (one group = the whole batch, or one image when per_image=True; in that
// case all images of a class are ordered by ONE sort of 64-bit keys with the image index in the high word):
//   lovasz_keys_kernel   p_c = ex2(z_c*log2e - lse*log2e) from the logits row of class c and the per-pixel log-sum-exp of
//                        ONE forward pass (b200seg_loss_fwd, WANT_LSE); error and foreground bit packed into one 32-bit
//                        sort key:  key = ((bits(e) + 1) << 1) | fg  (e >= 0, so its fp32 bit pattern is monotone;
//                        lossless); ignored pixels get key 0 and sink to the end of the descending order.
//   cub::DeviceRadixSort 4 digit passes over (key, pixel index) pairs — keys only when no gradient is wanted. A segment
//                        of a 512x1024x8 batch is 4 M pairs = 64 MB for both buffers of both arrays: L2 resident on B200,
//                        which is why segments are sorted one at a time instead of as one (class, error) 64-bit sort.
//                        (Library code: the CUDA toolkit's CUB, compiled into this .so; everything else is hand-written.)
//   lovasz_count_kernel / lovasz_tilescan_kernel / lovasz_grad_kernel
//                        exclusive scan of the foreground bits over the sorted order (tile counts -> one-CTA scan ->
//                        per-tile rescan), then with EXACT integer counts cum_i = #fg in [0,i], I_i = gts - cum_i,
//                        U_i = gts + (i + 1 - cum_i) the Jaccard increment in closed form
//                            g_i = 1/U_i              (fg_i = 1)
//                            g_i = I_i/(U_i (U_i-1))  (fg_i = 0),   g_0 = 1 - I_0/U_0
//                        (the reference's fp32 J_i - J_{i-1} cancels catastrophically: for P = 4 M its increments carry
//                        ~25 % noise; its cumsums stop being exact at 2^24), loss_c = sum e_i g_i, and
//                        dloss_c/dp_c = -+g_i scattered to the pixel's slot of G (C,N,H,W) f32 (class-major: a segment
//                        is one contiguous slice, the scatter index is the sorted value itself).
//   lovasz_finalize_kernel  mean over the present / all / listed classes, class weights, per-image reduction
//                        (weight_reduce_loss), and the coefficient table of the backward. No host sync anywhere.
//   lovasz_bwd_kernel    softmax Jacobian: grad_z_j = up * p_j (a_j - sum_c a_c p_c), a_c = coef_c G_c.
// The binary hinge variant (lovasz_hinge_flat :69-91) runs the same pipeline on errors 1 - z*sign with an order-preserving
// float->uint key and the foreground bit in the value's top bit.
//
// Ties: the loss is invariant to the order inside a block of equal errors (the block's increments telescope); the
// gradient is not (neither is the reference's: torch.sort's order among ties is unspecified).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace b200seg {

constexpr int kLovThreads = 256;
constexpr int kLovItems = 8;
constexpr int kLovTile = kLovThreads * kLovItems;   // sorted items per CTA in the scan kernels
constexpr int16_t kLovIgnored = -1;

__device__ __forceinline__ uint32_t hinge_key(float e) {   // order-preserving float -> uint, 0 reserved for "ignored"
  const uint32_t b = __float_as_uint(e);
  const uint32_t k = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return k < 1u ? 1u : k;
}
__device__ __forceinline__ float hinge_err(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// labels (any dtype) -> int16 class id: -1 = ignored, C = valid pixel that belongs to no class (label outside [0,C))
__global__ void __launch_bounds__(256) lovasz_prep_kernel(const void* __restrict__ labels, int label_dtype, long long n,
                                                          int C, int has_ignore, long long ignore,
                                                          int16_t* __restrict__ lab16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long y = load_label(labels, label_dtype, (size_t)i);
  int16_t r;
  if (has_ignore && y == ignore) r = kLovIgnored;
  else r = (y >= 0 && y < (long long)C) ? (int16_t)y : (int16_t)C;
  lab16[i] = r;
}

// ---------------------------------------------------------------------------------------------- sort keys
struct LovKeysParams {
  const void* logits;     // (N,C,HW) multi-class, (N,HW) binary
  const float* lse;       // (N,HW), multi-class only
  const int16_t* lab16;   // (N,HW)
  void* keys;             // (imgs * HW) uint32, or uint64 in batched mode
  uint32_t* vals;         // same count, or NULL (keys only)
  long long HW;
  int C, c;
  int n_img;              // number of images in the launch (batched mode: = segments)
};

// KeyT = uint32_t: one segment per launch (the images blockIdx.y of one group are concatenated, value = index in the group).
// KeyT = uint64_t: BATCHED per-image mode — every image is its own segment, all of them sorted by ONE radix sort: the
// high word carries (n_img - 1 - image) so that the descending order lists image 0 first; value = pixel index in the image.
template <typename T, int V, bool BINARY, typename KeyT>
__global__ void __launch_bounds__(256) lovasz_keys_kernel(const LovKeysParams p) {
  constexpr bool kBatched = sizeof(KeyT) == 8;
  const int nl = blockIdx.y;
  const long long hw0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (hw0 >= p.HW) return;
  const size_t n = (size_t)nl;
  const size_t px = n * p.HW + hw0;
  const T* zp = reinterpret_cast<const T*>(p.logits) + (BINARY ? px : (n * p.C + p.c) * p.HW + hw0);
  float z[V];
  load_vec<T, V>(zp, z);
  int lab[V];
  if constexpr (V == 4) {
    const uint2 r = ld_stream8(p.lab16 + px);
    lab[0] = (int16_t)(r.x & 0xffffu); lab[1] = (int16_t)(r.x >> 16);
    lab[2] = (int16_t)(r.y & 0xffffu); lab[3] = (int16_t)(r.y >> 16);
  } else {
    lab[0] = p.lab16[px];
  }
  uint32_t key[V], val[V];
  const size_t i0 = (size_t)nl * p.HW + hw0;                               // slot in the key / value arrays
  const uint32_t v0 = kBatched ? (uint32_t)hw0 : (uint32_t)i0;             // index inside the segment
  if constexpr (BINARY) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const uint32_t fg = lab[v] != 0 && lab[v] != kLovIgnored;       // labels are 0 / 1 (:78-79)
      const float sign = fg ? 1.f : -1.f;
      key[v] = lab[v] == kLovIgnored ? 0u : hinge_key(1.f - z[v] * sign);
      val[v] = (v0 + v) | (fg << 31);
    }
  } else {
    float l[V];
    load_vec<float, V>(p.lse + px, l);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float pc = ex2(fmaf(z[v], kLog2e, -l[v] * kLog2e));
      const uint32_t fg = lab[v] == p.c;
      const float e = fabsf((fg ? 1.f : 0.f) - pc);
      key[v] = lab[v] == kLovIgnored ? 0u : (((__float_as_uint(e) + 1u) << 1) | fg);
      val[v] = v0 + v;
    }
  }
  KeyT* kout = reinterpret_cast<KeyT*>(p.keys) + i0;
  if constexpr (kBatched) {
    const uint32_t hi = (uint32_t)(p.n_img - 1 - nl);
    if constexpr (V == 4) {
      reinterpret_cast<uint4*>(kout)[0] = make_uint4(key[0], hi, key[1], hi);
      reinterpret_cast<uint4*>(kout)[1] = make_uint4(key[2], hi, key[3], hi);
    } else {
      kout[0] = ((KeyT)hi << 32) | key[0];
    }
  } else {
    if constexpr (V == 4) *reinterpret_cast<uint4*>(kout) = make_uint4(key[0], key[1], key[2], key[3]);
    else kout[0] = key[0];
  }
  if (p.vals) {
    if constexpr (V == 4) *reinterpret_cast<uint4*>(p.vals + i0) = make_uint4(val[0], val[1], val[2], val[3]);
    else p.vals[i0] = val[0];
  }
}

// ---------------------------------------------------------------------------------------------- scan over the sorted order
// 8 consecutive sorted items of one segment (low word of the key = error | foreground bit; the high word of a batched
// 64-bit key is the image index, not needed once the order is established)
template <typename KeyT>
__device__ __forceinline__ void lov_load_tile(const KeyT* __restrict__ keys, const uint32_t* __restrict__ vals,
                                              long long i0, long long len, bool vec, uint32_t (&k)[kLovItems],
                                              uint32_t (&v)[kLovItems], bool want_vals) {
  if (vec && i0 + kLovItems <= len) {
    if constexpr (sizeof(KeyT) == 4) {
      const uint4 a = *reinterpret_cast<const uint4*>(keys + i0), b = *reinterpret_cast<const uint4*>(keys + i0 + 4);
      k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 a = *reinterpret_cast<const uint4*>(keys + i0 + 2 * j);
        k[2 * j] = a.x; k[2 * j + 1] = a.z;
      }
    }
    if (want_vals) {
      const uint4 c = *reinterpret_cast<const uint4*>(vals + i0), d = *reinterpret_cast<const uint4*>(vals + i0 + 4);
      v[0] = c.x; v[1] = c.y; v[2] = c.z; v[3] = c.w; v[4] = d.x; v[5] = d.y; v[6] = d.z; v[7] = d.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kLovItems; ++j) {
      const bool in = i0 + j < len;
      k[j] = in ? (uint32_t)keys[i0 + j] : 0u;
      v[j] = (in && want_vals) ? vals[i0 + j] : 0u;
    }
  }
}
template <bool BINARY> __device__ __forceinline__ uint32_t lov_fg(uint32_t key, uint32_t val) {
  if constexpr (BINARY) return key ? (val >> 31) : 0u;
  else return key & 1u;                                  // ignored items carry key 0
}

// grid (tiles per segment, segments): foreground count of every tile
template <bool BINARY, typename KeyT>
__global__ void __launch_bounds__(kLovThreads) lovasz_count_kernel(const KeyT* __restrict__ keys,
                                                                   const uint32_t* __restrict__ vals, long long len,
                                                                   int vec, uint32_t* __restrict__ tile_cnt) {
  __shared__ uint32_t s[kLovThreads / 32];
  const size_t seg0 = (size_t)blockIdx.y * len;
  const long long i0 = ((long long)blockIdx.x * kLovThreads + threadIdx.x) * kLovItems;
  uint32_t k[kLovItems], v[kLovItems];
  lov_load_tile<KeyT>(keys + seg0, vals ? vals + seg0 : nullptr, i0, len, vec != 0, k, v, BINARY);
  uint32_t cnt = 0;
#pragma unroll
  for (int j = 0; j < kLovItems; ++j) cnt += lov_fg<BINARY>(k[j], v[j]);
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kLovThreads / 32; ++w) t += s[w];
    tile_cnt[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

// one CTA per segment: exclusive scan of its tile counts; writes the segment's foreground total (+1 = "segment processed")
__global__ void __launch_bounds__(1024) lovasz_tilescan_kernel(const uint32_t* __restrict__ tile_cnt_all,
                                                               uint32_t* __restrict__ tile_off_all, int nb,
                                                               double* __restrict__ seg_stat_all, int seg_stat_stride) {
  const uint32_t* tile_cnt = tile_cnt_all + (size_t)blockIdx.x * nb;      // one CTA per segment
  uint32_t* tile_off = tile_off_all + (size_t)blockIdx.x * nb;
  double* seg_stat = seg_stat_all + (size_t)blockIdx.x * seg_stat_stride;
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nb ? tile_cnt[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_w[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    const uint32_t incl = carry + (warp ? s_w[warp - 1] : 0u) + x;
    if (i < nb) tile_off[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) seg_stat[1] = (double)s_carry + 1.0;
}

struct LovGradParams {
  const void* keys;
  const uint32_t* vals;       // NULL when no gradient is wanted (multi-class)
  const uint32_t* tile_off;   // (segments, tiles per segment)
  double* seg_stat;           // segment 0: [0] loss accumulator, [1] gts + 1; segment s at + s * seg_stat_stride
  float* G;                   // NULL = forward only
  float* Gseg;                // slice of G where segment 0 starts: multi-class (C,N,HW) f32 at class c, binary (N,HW)
  long long len;              // items per segment (consecutive segments are `len` apart in keys, vals and G)
  int seg_stat_stride;
  int vec;
};

template <bool BINARY, typename KeyT>
__global__ void __launch_bounds__(kLovThreads) lovasz_grad_kernel(const LovGradParams p) {
  __shared__ uint32_t s_w[kLovThreads / 32];
  __shared__ float s_l[kLovThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long i0 = ((long long)blockIdx.x * kLovThreads + threadIdx.x) * kLovItems;
  const bool want_vals = BINARY || p.G != nullptr;
  uint32_t k[kLovItems], v[kLovItems];
  const size_t seg0 = (size_t)blockIdx.y * p.len;
  double* seg_stat = p.seg_stat + (size_t)blockIdx.y * p.seg_stat_stride;
  lov_load_tile<KeyT>(reinterpret_cast<const KeyT*>(p.keys) + seg0, p.vals ? p.vals + seg0 : nullptr, i0, p.len, p.vec != 0, k, v,
                      want_vals);
  uint32_t tsum = 0;
#pragma unroll
  for (int j = 0; j < kLovItems; ++j) tsum += lov_fg<BINARY>(k[j], v[j]);
  // exclusive scan of the per-thread sums across the CTA
  uint32_t x = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_w[warp] = x;
  __syncthreads();
  uint32_t wpre = 0;
#pragma unroll
  for (int w = 0; w < kLovThreads / 32; ++w) wpre += (w < warp) ? s_w[w] : 0u;
  uint32_t cum = p.tile_off[(size_t)blockIdx.y * gridDim.x + blockIdx.x] + wpre + (x - tsum);
  const uint32_t gts = (uint32_t)(seg_stat[1] - 1.0);
  float loss = 0.f;
#pragma unroll
  for (int j = 0; j < kLovItems; ++j) {
    if (k[j] == 0u) continue;                         // ignored pixel / past the end
    const uint32_t fg = lov_fg<BINARY>(k[j], v[j]);
    cum += fg;
    const long long i = i0 + j;
    const float I = (float)(gts - cum);
    const float U = (float)((unsigned long long)gts + (unsigned long long)(i + 1) - cum);
    // closed-form Jaccard increment; rcp.approx (1 ulp) instead of IEEE divisions. i == 0: J_0 = 1 - I/U = (U - I)/U
    const float num = i == 0 ? U - I : (fg ? 1.f : I);
    const float den = (i == 0 || fg) ? U : U * (U - 1.f);
    const float g = num * fast_rcp(den);
    float e, dG;
    uint32_t idx;
    if constexpr (BINARY) {
      e = hinge_err(k[j]);
      idx = v[j] & 0x7fffffffu;
      dG = e > 0.f ? (fg ? -g : g) : 0.f;             // d relu(1 - z*sign)/dz = -sign
      e = fmaxf(e, 0.f);
    } else {
      e = __uint_as_float((k[j] >> 1) - 1u);
      idx = v[j];
      dG = fg ? -g : g;                               // d|fg - p|/dp
    }
    loss = fmaf(e, g, loss);
    if (p.G) {
      p.Gseg[seg0 + idx] = dG;                         // G is class-major: (C, N, HW), a segment is contiguous
    }
  }
  loss = warp_sum(loss);
  if (lane == 0) s_l[warp] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kLovThreads / 32; ++w) t += (double)s_l[w];
    if (t != 0.0) atomicAdd(seg_stat, t);
  }
}

// ---------------------------------------------------------------------------------------------- finalize
struct LovFinParams {
  const double* seg_stats;    // (n_groups, n_seg, 2)
  const float* cw;            // (n_seg) or NULL
  float* out;                 // n_groups floats (reduction none) or 1
  float* coef;                // (n_groups, n_seg) or NULL
  int n_groups, n_seg, only_present, per_image, reduction, has_avg_factor;
  double avg_factor;
  float loss_weight;
};

__global__ void __launch_bounds__(256) lovasz_finalize_kernel(const LovFinParams p) {
  __shared__ double sred[32];
  // scale of every group's loss inside the returned value (weight_reduce_loss, models/losses/utils.py:48-80)
  double gscale = 1.0;
  if (p.per_image && p.reduction == B200SEG_RED_MEAN)
    gscale = p.has_avg_factor ? 1.0 / (double)(float)((float)p.avg_factor + 1.1920928955078125e-07f) : 1.0 / (double)p.n_groups;
  double acc[1] = {0.0};
  for (int g = threadIdx.x; g < p.n_groups; g += blockDim.x) {
    const double* st = p.seg_stats + (size_t)g * p.n_seg * 2;
    int cnt = 0;
    double sum = 0.0;
    for (int c = 0; c < p.n_seg; ++c) {
      const double t = st[2 * c + 1];
      if (t == 0.0) continue;                                   // class not in the requested list
      if (p.only_present && t == 1.0) continue;                 // no foreground pixel (:153-154)
      ++cnt;
      sum += (p.cw ? (double)p.cw[c] : 1.0) * st[2 * c];
    }
    const double lg = cnt ? sum / (double)cnt : 0.0;            // torch.stack(losses).mean() (:169)
    if (p.coef) {
      for (int c = 0; c < p.n_seg; ++c) {
        const double t = st[2 * c + 1];
        const bool in = t != 0.0 && !(p.only_present && t == 1.0);
        p.coef[(size_t)g * p.n_seg + c] =
            in ? (float)((double)p.loss_weight * gscale * (p.cw ? (double)p.cw[c] : 1.0) / (double)cnt) : 0.f;
      }
    }
    if (p.per_image && p.reduction == B200SEG_RED_NONE) p.out[g] = (float)((double)p.loss_weight * lg);
    else acc[0] += lg;
  }
  if (!(p.per_image && p.reduction == B200SEG_RED_NONE)) {
    block_sum<double, 1>(acc, sred);
    if (threadIdx.x == 0) p.out[0] = (float)((double)p.loss_weight * gscale * acc[0]);
  }
}

// ---------------------------------------------------------------------------------------------- backward
struct LovBwdParams {
  const void* logits;
  const float* lse;
  const int16_t* lab16;
  const float* G;
  const float* coef;          // (n_groups, C) multi-class; (n_groups) binary
  const float* grad_out;      // device f32: scalar, or one per group (reduction 'none'), or NULL (= 1)
  void* grad;
  long long HW;
  int C, per_image, grad_per_group;
};

template <typename T, int V>
__global__ void __launch_bounds__(256) lovasz_bwd_kernel(const LovBwdParams p) {
  const int n = blockIdx.y;
  const long long hw0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (hw0 >= p.HW) return;
  const int g = p.per_image ? n : 0;
  const float up = p.grad_out ? p.grad_out[p.grad_per_group ? g : 0] : 1.f;
  const size_t px = (size_t)n * p.HW + hw0;
  float nl[V];
  load_vec<float, V>(p.lse + px, nl);
  bool live[V];
  if constexpr (V == 4) {
    const uint2 r = ld_stream8(p.lab16 + px);
    live[0] = (int16_t)(r.x & 0xffffu) != kLovIgnored; live[1] = (int16_t)(r.x >> 16) != kLovIgnored;
    live[2] = (int16_t)(r.y & 0xffffu) != kLovIgnored; live[3] = (int16_t)(r.y >> 16) != kLovIgnored;
  } else {
    live[0] = p.lab16[px] != kLovIgnored;
  }
#pragma unroll
  for (int v = 0; v < V; ++v) nl[v] = -nl[v] * kLog2e;
  const T* zrow = reinterpret_cast<const T*>(p.logits) + (size_t)n * p.C * p.HW + hw0;
  const float* grow = p.G + (size_t)n * p.HW + hw0;            // (C, N, HW): class stride N * HW
  const size_t gstride = (size_t)gridDim.y * p.HW;
  const float* cf = p.coef + (size_t)g * p.C;
  float dot[V];
#pragma unroll
  for (int v = 0; v < V; ++v) dot[v] = 0.f;
  for (int c = 0; c < p.C; ++c) {
    const float k = __ldg(cf + c);
    if (k == 0.f) continue;                           // class left out: its slots of G were never written
    float z[V], gg[V];
    load_vec<T, V>(zrow + (size_t)c * p.HW, z);
    load_vec<float, V>(grow + (size_t)c * gstride, gg);
#pragma unroll
    for (int v = 0; v < V; ++v) dot[v] = fmaf(k * gg[v], ex2(fmaf(z[v], kLog2e, nl[v])), dot[v]);
  }
  T* out = reinterpret_cast<T*>(p.grad) + (size_t)n * p.C * p.HW + hw0;
  for (int c = 0; c < p.C; ++c) {
    const float k = __ldg(cf + c);
    float z[V], gg[V], r[V];
    load_vec<T, V>(zrow + (size_t)c * p.HW, z);
    if (k != 0.f) {
      load_vec<float, V>(grow + (size_t)c * gstride, gg);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) gg[v] = 0.f;
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float pc = ex2(fmaf(z[v], kLog2e, nl[v]));
      r[v] = live[v] ? up * pc * (k * gg[v] - dot[v]) : 0.f;
    }
    store_vec<T, V>(out + (size_t)c * p.HW, r);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) lovasz_hinge_bwd_kernel(const LovBwdParams p) {
  const int n = blockIdx.y;
  const long long hw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (hw >= p.HW) return;
  const int g = p.per_image ? n : 0;
  const float up = (p.grad_out ? p.grad_out[p.grad_per_group ? g : 0] : 1.f) * p.coef[g];
  const size_t px = (size_t)n * p.HW + hw;
  const float r = p.lab16[px] != kLovIgnored ? up * p.G[px] : 0.f;
  reinterpret_cast<T*>(p.grad)[px] = from_float<T>(r);
}

// ---------------------------------------------------------------------------------------------- host side
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int bits_for(int n) {   // bits needed to hold values 0 .. n-1 (>= 1)
  int b = 1;
  while ((1 << b) < n) ++b;
  return b;
}

struct LovWorkspace {
  void *keys_a, *keys_b;
  uint32_t *vals_a, *vals_b, *tile_cnt, *tile_off;
  void* cub_temp;
  size_t cub_bytes, total;
};

// items = keys per sort; segs = segments per sort (> 1: batched per-image mode with 64-bit keys)
static int lov_carve(long long items, int segs, bool pairs, void* base, LovWorkspace* w) {
  size_t cub_bytes = 0;
  const int n = (int)items;
  cudaError_t e;
  if (segs > 1) {
    const int end_bit = 32 + bits_for(segs);
    e = pairs ? cub::DeviceRadixSort::SortPairsDescending(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                                          (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, end_bit, (cudaStream_t)0)
              : cub::DeviceRadixSort::SortKeysDescending(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, n, 0,
                                                         end_bit, (cudaStream_t)0);
  } else {
    e = pairs ? cub::DeviceRadixSort::SortPairsDescending(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                                          (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 32, (cudaStream_t)0)
              : cub::DeviceRadixSort::SortKeysDescending(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 32,
                                                         (cudaStream_t)0);
  }
  if (e != cudaSuccess) {
    set_error("lovasz: radix-sort workspace query failed: %s", cudaGetErrorString(e));
    return 2;
  }
  const size_t karr = align256((size_t)items * (segs > 1 ? 8 : 4));
  const size_t varr = align256((size_t)items * 4);
  const long long seg_len = items / segs;
  const size_t nb = (size_t)((seg_len + kLovTile - 1) / kLovTile) * segs;
  const size_t tb = align256((nb + 1) * sizeof(uint32_t));
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  w->keys_a = p + off; off += karr;
  w->keys_b = p + off; off += karr;
  w->vals_a = reinterpret_cast<uint32_t*>(p + off); off += pairs ? varr : 0;
  w->vals_b = reinterpret_cast<uint32_t*>(p + off); off += pairs ? varr : 0;
  w->tile_cnt = reinterpret_cast<uint32_t*>(p + off); off += tb;
  w->tile_off = reinterpret_cast<uint32_t*>(p + off); off += tb;
  w->cub_temp = p + off; off += align256(cub_bytes ? cub_bytes : 1);
  w->cub_bytes = cub_bytes;
  w->total = off;
  return 0;
}

long long lovasz_workspace_bytes(long long seg_len, int segs, int pairs) {
  if (seg_len <= 0 || segs <= 0) return 256;
  LovWorkspace w;
  if (lov_carve(seg_len * segs, segs, pairs != 0, nullptr, &w)) return -1;
  return (long long)w.total;
}

template <typename T, typename KeyT>
static int lov_fwd_typed(const b200seg_lovasz_desc* d, cudaStream_t st) {
  constexpr bool kBatched = sizeof(KeyT) == 8;
  const bool binary = d->binary != 0;
  const int n_seg = binary ? 1 : d->C;
  const long long HW = d->HW;
  // kBatched: every image is a segment and ONE sort orders all of them; else one segment = the whole batch (or N == 1)
  const int segs = kBatched ? d->N : 1;
  const long long seg_len = kBatched ? HW : (long long)d->N * HW;
  const long long items = seg_len * segs;
  const bool pairs = binary || d->G != nullptr;
  LovWorkspace w;
  if (int e = lov_carve(items, segs, pairs, d->workspace, &w)) return e;
  B200SEG_REQUIRE((long long)w.total <= d->workspace_bytes, "lovasz_fwd: workspace of %lld bytes needed, %lld given",
                  (long long)w.total, (long long)d->workspace_bytes);
  const long long npx = (long long)d->N * HW;
  lovasz_prep_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(d->labels, d->label_dtype, npx, binary ? 2 : d->C,
                                                                    d->has_ignore, d->ignore_index, d->lab16);
  count_launch();
  if (int e = check_launch("lovasz_prep_kernel")) return e;

  const bool vec = HW % 4 == 0 && aligned16(d->logits) && aligned16(d->lab16) && (binary || aligned16(d->lse));
  const int scan_vec = (segs == 1 || seg_len % 8 == 0) ? 1 : 0;      // 16-byte loads of the sorted arrays stay aligned
  const int nb = (int)((seg_len + kLovTile - 1) / kLovTile);
  const int end_bit = kBatched ? 32 + bits_for(segs) : 32;
  const int n_list = binary ? 1 : (d->classes_host ? d->n_classes : d->C);
  for (int j = 0; j < n_list; ++j) {
    const int c = binary ? 0 : (d->classes_host ? d->classes_host[j] : j);
    LovKeysParams kp;
    kp.logits = d->logits; kp.lse = d->lse; kp.lab16 = d->lab16;
    kp.keys = w.keys_a; kp.vals = pairs ? w.vals_a : nullptr;
    kp.HW = HW; kp.C = d->C; kp.c = c; kp.n_img = d->N;
    if (vec) {
      dim3 grid((unsigned)((HW / 4 + 255) / 256), d->N);
      if (binary) lovasz_keys_kernel<T, 4, true, KeyT><<<grid, 256, 0, st>>>(kp);
      else lovasz_keys_kernel<T, 4, false, KeyT><<<grid, 256, 0, st>>>(kp);
    } else {
      dim3 grid((unsigned)((HW + 255) / 256), d->N);
      if (binary) lovasz_keys_kernel<T, 1, true, KeyT><<<grid, 256, 0, st>>>(kp);
      else lovasz_keys_kernel<T, 1, false, KeyT><<<grid, 256, 0, st>>>(kp);
    }
    if (int e = check_launch("lovasz_keys_kernel")) return e;
    size_t cb = w.cub_bytes;
    const KeyT* kin = reinterpret_cast<const KeyT*>(w.keys_a);
    KeyT* kout = reinterpret_cast<KeyT*>(w.keys_b);
    cudaError_t ce = pairs ? cub::DeviceRadixSort::SortPairsDescending(w.cub_temp, cb, kin, kout, (const uint32_t*)w.vals_a,
                                                                       w.vals_b, (int)items, 0, end_bit, st)
                           : cub::DeviceRadixSort::SortKeysDescending(w.cub_temp, cb, kin, kout, (int)items, 0, end_bit, st);
    if (ce != cudaSuccess) {
      set_error("lovasz_fwd: radix sort failed: %s", cudaGetErrorString(ce));
      return 2;
    }
    // segment s of this launch = image group s: statistics at (s, c), G slice at (c, s)
    double* seg = d->seg_stats + (size_t)c * 2;
    const int seg_stride = n_seg * 2;
    dim3 sgrid(nb, segs);
    const uint32_t* vsorted = pairs ? w.vals_b : nullptr;
    if (binary) lovasz_count_kernel<true, KeyT><<<sgrid, kLovThreads, 0, st>>>(kout, vsorted, seg_len, scan_vec, w.tile_cnt);
    else lovasz_count_kernel<false, KeyT><<<sgrid, kLovThreads, 0, st>>>(kout, nullptr, seg_len, scan_vec, w.tile_cnt);
    lovasz_tilescan_kernel<<<segs, 1024, 0, st>>>(w.tile_cnt, w.tile_off, nb, seg, seg_stride);
    LovGradParams gp;
    gp.keys = kout; gp.vals = vsorted; gp.tile_off = w.tile_off; gp.seg_stat = seg; gp.seg_stat_stride = seg_stride;
    gp.G = d->G; gp.len = seg_len; gp.vec = scan_vec;
    gp.Gseg = d->G ? d->G + (size_t)(binary ? 0 : c) * d->N * HW : nullptr;
    if (binary) lovasz_grad_kernel<true, KeyT><<<sgrid, kLovThreads, 0, st>>>(gp);
    else lovasz_grad_kernel<false, KeyT><<<sgrid, kLovThreads, 0, st>>>(gp);
    count_launch(4);
    if (int e = check_launch("lovasz scan kernels")) return e;
  }
  return 0;
}

int lovasz_fwd_dispatch(const b200seg_lovasz_desc* d, cudaStream_t st) {
  const int n_seg = d->binary ? 1 : d->C;
  const int n_groups = d->per_image ? d->N : 1;
  B200SEG_CUDA(cudaMemsetAsync(d->seg_stats, 0, (size_t)n_groups * n_seg * 2 * sizeof(double), st));
  if (d->N > 0 && d->HW > 0) {
    int e;
    const bool batched = d->per_image && d->N > 1;   // all images of a class in one 64-bit-key sort
    switch (d->logit_dtype) {
      case B200SEG_F32: e = batched ? lov_fwd_typed<float, uint64_t>(d, st) : lov_fwd_typed<float, uint32_t>(d, st); break;
      case B200SEG_BF16:
        e = batched ? lov_fwd_typed<__nv_bfloat16, uint64_t>(d, st) : lov_fwd_typed<__nv_bfloat16, uint32_t>(d, st);
        break;
      default: e = batched ? lov_fwd_typed<__half, uint64_t>(d, st) : lov_fwd_typed<__half, uint32_t>(d, st); break;
    }
    if (e) return e;
  }
  LovFinParams fp;
  fp.seg_stats = d->seg_stats; fp.cw = d->binary ? nullptr : d->class_weight; fp.out = d->out; fp.coef = d->coef;
  fp.n_groups = n_groups; fp.n_seg = n_seg; fp.only_present = d->binary ? 0 : d->only_present;
  fp.per_image = d->per_image; fp.reduction = d->reduction; fp.has_avg_factor = d->has_avg_factor;
  fp.avg_factor = d->avg_factor; fp.loss_weight = d->loss_weight;
  if (n_groups == 0) {   // empty batch: weight_reduce_loss of an empty stack
    B200SEG_CUDA(cudaMemsetAsync(d->out, 0, sizeof(float), st));
    return 0;
  }
  lovasz_finalize_kernel<<<1, 256, 0, st>>>(fp);
  count_launch();
  return check_launch("lovasz_finalize_kernel");
}

template <typename T> static int lov_bwd_typed(const b200seg_lovasz_bwd_desc* d, cudaStream_t st) {
  LovBwdParams p;
  p.logits = d->logits; p.lse = d->lse; p.lab16 = d->lab16; p.G = d->G; p.coef = d->coef; p.grad_out = d->grad_out;
  p.grad = d->grad_logits; p.HW = d->HW; p.C = d->C; p.per_image = d->per_image; p.grad_per_group = d->grad_per_group;
  if (d->binary) {
    dim3 grid((unsigned)((d->HW + 255) / 256), d->N);
    lovasz_hinge_bwd_kernel<T><<<grid, 256, 0, st>>>(p);
  } else {
    const bool vec = d->HW % 4 == 0 && aligned16(d->logits) && aligned16(d->lab16) && aligned16(d->lse) &&
                     aligned16(d->G) && aligned16(d->grad_logits);
    if (vec) {
      dim3 grid((unsigned)((d->HW / 4 + 255) / 256), d->N);
      lovasz_bwd_kernel<T, 4><<<grid, 256, 0, st>>>(p);
    } else {
      dim3 grid((unsigned)((d->HW + 255) / 256), d->N);
      lovasz_bwd_kernel<T, 1><<<grid, 256, 0, st>>>(p);
    }
  }
  count_launch();
  return check_launch("lovasz_bwd_kernel");
}

int lovasz_bwd_dispatch(const b200seg_lovasz_bwd_desc* d, cudaStream_t st) {
  switch (d->logit_dtype) {
    case B200SEG_F32: return lov_bwd_typed<float>(d, st);
    case B200SEG_BF16: return lov_bwd_typed<__nv_bfloat16>(d, st);
    default: return lov_bwd_typed<__half>(d, st);
  }
}

}  // namespace b200seg
