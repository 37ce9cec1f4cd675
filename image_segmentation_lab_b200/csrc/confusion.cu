// Fused arg-max + per-class area histograms (intersect / pred / label), sm_100a.
//
// Replaces SegEvaluator.intersect_and_union (core/evaluation/metrics.py:210-270) — three boolean compactions, three
// float casts, three torch.histc and three .cpu() syncs PER IMAGE — and, in the logits variant, the softmax+argmax
// of SegEvaluator.process (:101-107), with ONE launch for a whole list of images and no host synchronisation.
//
// Semantics kept (file:line): pixels with gt == ignore_index are dropped from all three histograms (:237-241);
// intersect counts pred where pred == gt (:248); torch.histc(bins=C, min=0, max=C-1) drops values outside [0, C-1]
// (:249-265), so an out-of-range gt still leaves its pixel in the pred histogram and vice versa. Areas are exact
// int64 (the reference stores them in fp32).
//
// Work split: the images are cut into chunks of kChunk pixels; persistent CTAs (SMs x resident CTAs) each take one
// contiguous range of chunks, so a CTA touches few images and flushes its counters to global memory once per image
// it touches. Every sample is decoded to a 32-bit class index (>= 0 in range, -1 out of range, -2 ignored) with
// dtype-specialised integer / float compares — no 64-bit arithmetic in the pixel loop.
// Counters: for small C every thread owns a private column cnt[bin][tid] in shared memory (bank == tid: plain
// conflict-free LDS/IADD/STS, no atomics — random predictions would otherwise serialise on ATOMS throughput); a
// pixel needs one update when pred == gt (bin A) and two otherwise (B = pred only, D = gt only): I = A, P = A + B,
// L = A + D. For large C a shared-memory atomic histogram with per-thread run-length aggregation is used.
//
// Roofline: HBM. Algorithmic bytes per pixel: pred bytes + gt bytes (label maps: 8 + 4 = 12; logits: C*s + 4).
#include <cstdlib>

#include "common.cuh"

namespace b200seg {

constexpr int kChunk = 4096;
constexpr int kIgnored = -2;
constexpr long long kMaxChunksPerFlush = 1ll << 18;   // totals mode: 2^30 pixels per CTA between flushes keep the 32-bit counters exact

struct ConfParams {
  const b200seg_image* images;
  const long long* chunk_prefix;
  int n_images;
  long long total_chunks;
  int pred_dtype, gt_dtype;
  int C;
  long long ignore;
  long long* areas;
  long long* const* pred_out;
  int totals_only;   // areas is (3,C): the sum over the images, flushed once per CTA instead of once per image touched
  float simple_ratio;   // resize-fused variant: images whose column ratio W / w is below this take the per-pixel form
};

// Decoder of V consecutive samples of a label-like tensor into class indices.
struct ClassDecoder {
  int dt, C;
  bool has_ignore;
  unsigned ign_lo, ign_hi;   // 64-bit pattern of ignore_index
  float ign_f;               // ignore_index as float (labels stored as float compare in float, as the reference)
  bool ign_fits_i32, ign_fits_u8;
  int ign_i32;

  __device__ __forceinline__ void init(int dtype, int classes, bool with_ignore, long long ignore) {
    dt = dtype; C = classes; has_ignore = with_ignore;
    ign_lo = (unsigned)((unsigned long long)ignore & 0xffffffffull);
    ign_hi = (unsigned)((unsigned long long)ignore >> 32);
    ign_f = (float)ignore;
    ign_fits_i32 = (ignore >= -2147483648ll && ignore <= 2147483647ll);
    ign_fits_u8 = (ignore >= 0 && ignore <= 255);
    ign_i32 = (int)ignore;
  }
  __device__ __forceinline__ int from_i64(unsigned lo, unsigned hi) const {
    if (has_ignore && lo == ign_lo && hi == ign_hi) return kIgnored;
    return (hi == 0u && lo < (unsigned)C) ? (int)lo : -1;
  }
  __device__ __forceinline__ int from_i32(int v) const {
    if (has_ignore && ign_fits_i32 && v == ign_i32) return kIgnored;
    return ((unsigned)v < (unsigned)C) ? v : -1;
  }
  __device__ __forceinline__ int from_f32(float f) const {
    if (has_ignore && f == ign_f) return kIgnored;
    return (f >= 0.f && f < (float)C) ? (int)f : -1;   // NaN -> -1
  }
  __device__ __forceinline__ int from_generic(long long v) const {
    if (has_ignore && v == (long long)(((unsigned long long)ign_hi << 32) | ign_lo)) return kIgnored;
    return (v >= 0 && v < (long long)C) ? (int)v : -1;
  }
  // scalar path (any dtype, any alignment)
  __device__ __forceinline__ int one(const void* p, size_t i) const {
    switch (dt) {
      case B200SEG_L_I64: { const uint2 r = reinterpret_cast<const uint2*>(p)[i]; return from_i64(r.x, r.y); }
      case B200SEG_L_F32: return from_f32(reinterpret_cast<const float*>(p)[i]);
      case B200SEG_L_I32: return from_i32(reinterpret_cast<const int*>(p)[i]);
      case B200SEG_L_U8: return from_i32((int)reinterpret_cast<const uint8_t*>(p)[i]);
      case B200SEG_L_I16: return from_i32((int)reinterpret_cast<const int16_t*>(p)[i]);
      default: {
        const double d = reinterpret_cast<const double*>(p)[i];
        if (has_ignore && d == (double)(long long)(((unsigned long long)ign_hi << 32) | ign_lo)) return kIgnored;
        return (d >= 0.0 && d < (double)C) ? (int)d : -1;
      }
    }
  }
  // up to N consecutive samples starting at element i (any alignment); slots >= n read as ignored. The dtype switch
  // is taken once, the loads of a case are independent of each other.
  template <int N> __device__ __forceinline__ void upto(const void* p, size_t i, int n, int (&o)[N]) const {
    if (dt == B200SEG_L_F32) {
      const float* q = reinterpret_cast<const float*>(p) + i;
      float v[N];
#pragma unroll
      for (int k = 0; k < N; ++k) v[k] = k < n ? __ldg(q + k) : 0.f;
#pragma unroll
      for (int k = 0; k < N; ++k) o[k] = k < n ? from_f32(v[k]) : kIgnored;
    } else if (dt == B200SEG_L_I64) {
      const uint2* q = reinterpret_cast<const uint2*>(p) + i;
      uint2 v[N];
#pragma unroll
      for (int k = 0; k < N; ++k) v[k] = k < n ? __ldg(q + k) : make_uint2(0u, 0u);
#pragma unroll
      for (int k = 0; k < N; ++k) o[k] = k < n ? from_i64(v[k].x, v[k].y) : kIgnored;
    } else if (dt == B200SEG_L_U8) {
      const uint8_t* q = reinterpret_cast<const uint8_t*>(p) + i;
      int v[N];
#pragma unroll
      for (int k = 0; k < N; ++k) v[k] = k < n ? (int)__ldg(q + k) : 0;
#pragma unroll
      for (int k = 0; k < N; ++k) o[k] = k < n ? from_i32(v[k]) : kIgnored;
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) o[k] = k < n ? one(p, i + k) : kIgnored;
    }
  }
  // 8 consecutive samples starting at element i (i % 8 == 0, base 16-byte aligned)
  __device__ __forceinline__ void eight(const void* p, size_t i, int (&o)[8]) const {
    const char* b = reinterpret_cast<const char*>(p);
    if (dt == B200SEG_L_I64) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 r = ld_stream16(b + i * 8 + 16 * k);
        o[2 * k] = from_i64(r.x, r.y);
        o[2 * k + 1] = from_i64(r.z, r.w);
      }
    } else if (dt == B200SEG_L_F32) {
      const uint4 a = ld_stream16(b + i * 4), c = ld_stream16(b + i * 4 + 16);
      o[0] = from_f32(__uint_as_float(a.x)); o[1] = from_f32(__uint_as_float(a.y));
      o[2] = from_f32(__uint_as_float(a.z)); o[3] = from_f32(__uint_as_float(a.w));
      o[4] = from_f32(__uint_as_float(c.x)); o[5] = from_f32(__uint_as_float(c.y));
      o[6] = from_f32(__uint_as_float(c.z)); o[7] = from_f32(__uint_as_float(c.w));
    } else if (dt == B200SEG_L_I32) {
      const uint4 a = ld_stream16(b + i * 4), c = ld_stream16(b + i * 4 + 16);
      o[0] = from_i32((int)a.x); o[1] = from_i32((int)a.y); o[2] = from_i32((int)a.z); o[3] = from_i32((int)a.w);
      o[4] = from_i32((int)c.x); o[5] = from_i32((int)c.y); o[6] = from_i32((int)c.z); o[7] = from_i32((int)c.w);
    } else if (dt == B200SEG_L_U8) {
      const uint2 a = ld_stream8(b + i);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = from_i32((int)(((k < 4 ? a.x : a.y) >> (8 * (k & 3))) & 0xffu));
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = one(p, i + k);
    }
  }
};

template <int THREADS, bool PRIVATE> struct Counters {
  unsigned int* cnt;
  int C;
  __device__ __forceinline__ void add_n(int bin, unsigned n) {
    if constexpr (PRIVATE) cnt[bin * THREADS + threadIdx.x] += n;
    else atomicAdd(cnt + bin, n);
  }
  // pv / gv: class index or -1 when outside [0, C). Atomic flavour (few, aggregated updates): plain branches.
  __device__ __forceinline__ void update(int pv, int gv, unsigned n = 1u) {
    if (pv == gv) {
      if (pv >= 0) add_n(pv, n);
    } else {
      if (pv >= 0) add_n(C + pv, n);
      if (gv >= 0) add_n(2 * C + gv, n);
    }
  }
  // Private flavour, one pixel: BRANCH-FREE — two unconditional read-modify-writes of this thread's own column
  // with a 0/1 increment (random predictions make every per-pixel branch divergent).
  //   match            : cnt[A + pv] += 1 ; cnt[D + gv] += 0
  //   mismatch         : cnt[B + pv] += 1 ; cnt[D + gv] += 1      (each only if its index is in range)
  //   gv == kIgnored   : both increments are 0
  __device__ __forceinline__ void update_private(unsigned int* mine, int pv, int gv) {
    const bool pin = pv >= 0, gin = gv >= 0;
    const bool match = pin & (pv == gv);
    const bool live = gv != kIgnored;
    const int b1 = pin ? (match ? pv : C + pv) : 0;
    const int b2 = gin ? 2 * C + gv : 0;
    const unsigned i1 = (live & pin) ? 1u : 0u;
    const unsigned i2 = (live & gin & !match) ? 1u : 0u;
    mine[b1 * THREADS] += i1;
    mine[b2 * THREADS] += i2;
  }
  __device__ __forceinline__ void zero() {
    const int total = PRIVATE ? 3 * C * THREADS : 3 * C;
    for (int i = threadIdx.x; i < total; i += THREADS) cnt[i] = 0u;
  }
  // add this CTA's counters into areas[img] = (I[C], P[C], L[C]) and clear them
  __device__ __forceinline__ void flush(long long* areas_img) {
    __syncthreads();
    for (int bin = threadIdx.x; bin < 3 * C; bin += THREADS) {
      unsigned long long tot = 0;
      if constexpr (PRIVATE) {
        for (int t = 0; t < THREADS; ++t) {
          const int tt = (t + threadIdx.x) & (THREADS - 1);  // rotate: conflict-free across the warp
          tot += cnt[bin * THREADS + tt];
          cnt[bin * THREADS + tt] = 0u;
        }
      } else {
        tot = cnt[bin];
        cnt[bin] = 0u;
      }
      if (tot) {
        unsigned long long* a = reinterpret_cast<unsigned long long*>(areas_img);
        const int kind = bin / C, c = bin - kind * C;
        if (kind == 0) {  // A: matched
          atomicAdd(a + c, tot);
          atomicAdd(a + C + c, tot);
          atomicAdd(a + 2 * C + c, tot);
        } else if (kind == 1) {  // B: pred only
          atomicAdd(a + C + c, tot);
        } else {  // D: gt only
          atomicAdd(a + 2 * C + c, tot);
        }
      }
    }
    __syncthreads();
  }
};

__device__ __forceinline__ int find_image(const long long* prefix, int n_images, long long chunk) {
  int lo = 0, hi = n_images - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (prefix[mid] <= chunk) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

template <typename T, int THREADS, bool PRIVATE, bool FROM_LOGITS>
__global__ void __launch_bounds__(THREADS) confusion_kernel(const ConfParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Counters<THREADS, PRIVATE> ctr{reinterpret_cast<unsigned int*>(smem_raw), p.C};
  ctr.zero();
  __syncthreads();
  const int C = p.C;
  const long long per = (p.total_chunks + gridDim.x - 1) / gridDim.x;
  long long chunk = (long long)blockIdx.x * per;
  const long long chunk_end = (chunk + per < p.total_chunks) ? chunk + per : p.total_chunks;
  if (chunk >= chunk_end) return;
  int img = find_image(p.chunk_prefix, p.n_images, chunk);
  long long since_flush = 0;
  constexpr int V = 8;
  ClassDecoder dgt, dpr;
  dgt.init(p.gt_dtype, C, true, p.ignore);
  dpr.init(p.pred_dtype, C, false, 0);

  while (chunk < chunk_end) {
    while (img + 1 < p.n_images && chunk >= p.chunk_prefix[img + 1]) ++img;  // skips empty images
    const b200seg_image im = p.images[img];
    const long long img_chunk_end = p.chunk_prefix[img + 1] < chunk_end ? p.chunk_prefix[img + 1] : chunk_end;
    const long long px_begin = (chunk - p.chunk_prefix[img]) * kChunk;
    long long px_end = (img_chunk_end - p.chunk_prefix[img]) * kChunk;
    if (px_end > im.n_pixels) px_end = im.n_pixels;
    const bool vec_ok = aligned16(im.pred) && aligned16(im.gt) && (!FROM_LOGITS || (im.n_pixels % V == 0));
    long long* pout = (FROM_LOGITS && p.pred_out) ? p.pred_out[img] : nullptr;

    long long px_scalar_from = px_begin;
    if constexpr (!FROM_LOGITS && PRIVATE) {
      // Fast path (label maps: int64 predictions + float32 ground truth, the reference's dtypes): software-pipelined
      // — the 6 x 128-bit loads of the NEXT iteration are issued before the current 8 pixels are decoded and counted.
      if (vec_ok && p.pred_dtype == B200SEG_L_I64 && p.gt_dtype == B200SEG_L_F32) {
        const long long span = (long long)THREADS * V;
        const long long n_full = (px_end - px_begin) / span;       // iterations in which every thread has 8 pixels
        const char* pb = reinterpret_cast<const char*>(im.pred);
        const char* gb = reinterpret_cast<const char*>(im.gt);
        unsigned int* mine = ctr.cnt + threadIdx.x;
        uint4 pr[4], gr[2];
        long long px = px_begin + (long long)threadIdx.x * V;
        if (n_full > 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) pr[k] = ld_stream16(pb + px * 8 + 16 * k);
          gr[0] = ld_stream16(gb + px * 4);
          gr[1] = ld_stream16(gb + px * 4 + 16);
        }
        for (long long itn = 0; itn < n_full; ++itn) {
          uint4 pn[4], gn[2];
          const long long pxn = px + span;
          if (itn + 1 < n_full) {
#pragma unroll
            for (int k = 0; k < 4; ++k) pn[k] = ld_stream16(pb + pxn * 8 + 16 * k);
            gn[0] = ld_stream16(gb + pxn * 4);
            gn[1] = ld_stream16(gb + pxn * 4 + 16);
          }
          // ... and the lines of the iteration after that are requested into L2: bytes in flight without registers
          // (measured: 0.90 -> 0.98 of the HBM copy peak; the register double buffer alone leaves 49 KB in flight per SM)
          if (itn + 2 < n_full) {
            const long long pxf = px + 2 * span;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + pxf * 8));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gb + pxf * 4));
          }
          const float gf[8] = {__uint_as_float(gr[0].x), __uint_as_float(gr[0].y), __uint_as_float(gr[0].z),
                               __uint_as_float(gr[0].w), __uint_as_float(gr[1].x), __uint_as_float(gr[1].y),
                               __uint_as_float(gr[1].z), __uint_as_float(gr[1].w)};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            ctr.update_private(mine, dpr.from_i64(pr[k].x, pr[k].y), dgt.from_f32(gf[2 * k]));
            ctr.update_private(mine, dpr.from_i64(pr[k].z, pr[k].w), dgt.from_f32(gf[2 * k + 1]));
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) pr[k] = pn[k];
          gr[0] = gn[0];
          gr[1] = gn[1];
          px = pxn;
        }
        px_scalar_from = px_begin + n_full * span;   // the ragged remainder goes through the generic loop below
      }
    }

    for (long long px = px_scalar_from + (long long)threadIdx.x * V; px < px_end; px += (long long)THREADS * V) {
      int gv[V], pv[V];
      const int nv = (px_end - px >= V) ? V : (int)(px_end - px);
      if (nv == V && vec_ok) {
        dgt.eight(im.gt, (size_t)px, gv);
        if constexpr (FROM_LOGITS) {
          float best[V];
#pragma unroll
          for (int v = 0; v < V; ++v) { best[v] = neg_inf(); pv[v] = 0; }
          const T* base = reinterpret_cast<const T*>(im.pred) + px;
          constexpr int LV = 16 / (int)sizeof(T);  // elements per 128-bit load
          for (int c = 0; c < C; ++c) {
            float z[V];
#pragma unroll
            for (int q = 0; q < V / LV; ++q) {
              float t[LV];
              load_vec<T, LV>(base + (size_t)c * im.n_pixels + q * LV, t);
#pragma unroll
              for (int k = 0; k < LV; ++k) z[q * LV + k] = t[k];
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
              if (z[v] > best[v]) { best[v] = z[v]; pv[v] = c; }  // lowest index wins ties
            }
          }
          if (pout) {
#pragma unroll
            for (int v = 0; v < V; v += 2)
              st_stream16(pout + px + v, make_uint4((unsigned)pv[v], 0u, (unsigned)pv[v + 1], 0u));
          }
        } else {
          dpr.eight(im.pred, (size_t)px, pv);
        }
      } else {
        for (int v = 0; v < V; ++v) {
          gv[v] = kIgnored;
          pv[v] = -1;
          if (v < nv) {
            gv[v] = dgt.one(im.gt, (size_t)(px + v));
            if constexpr (FROM_LOGITS) {
              const T* base = reinterpret_cast<const T*>(im.pred) + px + v;
              float best = neg_inf();
              int bi = 0;
              for (int c = 0; c < C; ++c) {
                const float z = to_float<T>(base[(size_t)c * im.n_pixels]);
                if (z > best) { best = z; bi = c; }
              }
              pv[v] = bi;
              if (pout) pout[px + v] = bi;
            } else {
              pv[v] = dpr.one(im.pred, (size_t)(px + v));
            }
          }
        }
      }
      if constexpr (PRIVATE) {
        unsigned int* mine = ctr.cnt + threadIdx.x;
#pragma unroll
        for (int v = 0; v < V; ++v) ctr.update_private(mine, pv[v], gv[v]);
      } else {
        // run-length aggregation over the thread's V consecutive pixels, then shared atomics
        int rp = -3, rg = -3;
        unsigned rn = 0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const bool live = gv[v] != kIgnored;
          if (live && pv[v] == rp && gv[v] == rg) {
            ++rn;
          } else {
            if (rn) ctr.update(rp, rg, rn);
            rn = live ? 1u : 0u;
            rp = live ? pv[v] : -3;
            rg = live ? gv[v] : -3;
          }
        }
        if (rn) ctr.update(rp, rg, rn);
      }
    }
    if (!p.totals_only) {
      ctr.flush(p.areas + (size_t)img * 3 * C);
    } else {
      since_flush += img_chunk_end - chunk;
      if (since_flush >= kMaxChunksPerFlush) { ctr.flush(p.areas); since_flush = 0; }
    }
    chunk = img_chunk_end;
  }
  if (p.totals_only) ctr.flush(p.areas);
}

// ------------------------------------------------------------------------------------------------
// Resize-fused variant: logits (C,h,w) are bilinearly interpolated to the ground-truth size (H,W) inside the
// arg-max loop (decode_head.py:297-320 rescale + metrics.py:101-107 argmax), so the (1,C,H,W) rescaled logits are
// never written. The interpolation is ATen's, operation for operation with its FMA contraction pinned (common.cuh:
// aten_src_index / aten_bilerp), and for 16-bit logits the value is rounded to the logit dtype before the comparison,
// as F.interpolate's output is: the arg-max — and with it every area — is BIT-EXACT against the reference's
// resize -> argmax at any ratio and either align_corners setting (on inputs without soft-max rounding ties).
//
// Work unit = a BAND of up to 4 output rows sharing the same pair of source rows x one RUN of columns sharing the same pair
// of source columns (resize_band_units below; rows that are not up-sampled keep the one-row form of the kernel body): a
// thread loads the 4 taps of a class once per unit (the next class's are in flight meanwhile), forms the horizontal sums
// once per column and evaluates FMUL + FFMA + compare / select per class-pixel, in chunks of 8 / 4 / 2 columns. Band and
// run boundaries come from the source-index maps themselves (any ratio), tabulated per image in shared memory. Bound:
// instruction issue / the ALU pipe (C interpolations per output pixel); DRAM traffic is the ground-truth map (4 B per
// pixel) plus the small logits.
constexpr float kSimpleRatio = 1.5f;   // (see ConfParams::simple_ratio; measured: the forms meet at a ratio of 1.7, down-sampling is 3 x faster per pixel)
constexpr int kRunTableMax = 4096;   // w + 2 entries per image; wider logits take the simple per-pixel kernel

// band key of an output position: 0 = source index clamped to 0 (align_corners=False), k + 1 = source floor k
__device__ __forceinline__ int aten_run_key(float scale, int dst, int in, bool ac) {
  const float raw = ac ? __fmul_rn(scale, (float)dst) : __fmaf_rn(scale, __fadd_rn((float)dst, 0.5f), -0.5f);
  if (raw < 0.f) return 0;
  const int i = (int)raw;
  return (i < in - 1 ? i : in - 1) + 1;
}
// first output position in [0, out] whose key is >= r (keys are non-decreasing)
__device__ __forceinline__ int aten_run_start(float scale, int r, int in, int out, bool ac) {
  if (r <= 0) return 0;
  if (r > in || !(scale > 0.f)) return out;
  const float est = ac ? ((float)(r - 1) / scale) : (((float)(r - 1) + 0.5f) / scale - 0.5f);
  int d = (int)fminf(fmaxf(ceilf(est), 0.f), (float)out);
  while (d > 0 && aten_run_key(scale, d - 1, in, ac) >= r) --d;
  while (d < out && aten_run_key(scale, d, in, ac) < r) ++d;
  return d;
}

// Per-pixel form (no run table) for images whose logits are wider than the table: one pixel per thread, 4 taps per class.
template <typename T, int THREADS, bool PRIVATE>
__device__ __forceinline__ void resize_pixels_simple(const b200seg_image& im, long long px_begin, long long px_end, float sh, float sw,
                                                     bool ac, int C, const ClassDecoder& dgt, long long* pout,
                                                     Counters<THREADS, PRIVATE>& ctr) {
  const long long hw = (long long)im.h * im.w;
  const T* base = reinterpret_cast<const T*>(im.pred);
  for (long long px = px_begin + threadIdx.x; px < px_end; px += THREADS) {
    const int gv = dgt.one(im.gt, (size_t)px);
    const int Y = (int)(px / im.W), X = (int)(px - (long long)Y * im.W);
    int y0, y1, x0, x1;
    float ly, lx;
    resize_src(sh, Y, im.h, ac, y0, y1, ly);
    resize_src(sw, X, im.w, ac, x0, x1, lx);
    const float h1 = ly, h0 = __fsub_rn(1.f, ly), w1 = lx, w0 = __fsub_rn(1.f, lx);
    const int o00 = y0 * im.w + x0, o01 = y0 * im.w + x1, o10 = y1 * im.w + x0, o11 = y1 * im.w + x1;
    float best = neg_inf();
    int bi = 0;
    const T* pl = base;
    for (int c = 0; c < C; ++c) {
      float z = aten_bilerp(h0, h1, w0, w1, to_float<T>(pl[o00]), to_float<T>(pl[o01]), to_float<T>(pl[o10]), to_float<T>(pl[o11]));
      if constexpr (sizeof(T) == 2) z = to_float<T>(from_float<T>(z));
      if (z > best) { best = z; bi = c; }   // lowest index wins ties
      pl += hw;
    }
    if (pout) pout[px] = bi;
    if constexpr (PRIVATE) ctr.update_private(ctr.cnt + threadIdx.x, bi, gv);
    else if (gv != kIgnored) ctr.update(bi, gv);
  }
}

// Band form of the (row, run) unit for UP-SAMPLED rows: R output rows that share the same pair of source rows (a "band" of
// the row index map, exactly like a run of the column map) are taken by one thread together. The horizontal sums of
// ATen's expression, X = fma(w0, v00, w1 v01) and Y = fma(w0, v10, w1 v11), depend on the source rows and the output
// column only, so they are formed ONCE per class and column and shared by the R rows; a row then costs FMUL + FFMA and
// the compare / select triple: 6.4 issue slots per class-pixel at R = 4 against 10.9 for the row form, with the
// per-unit work (index decomposition, weights, tap addresses) spread over 4 x as many pixels. Same operations on the
// same operands as aten_bilerp: bit-exact. The R rows' class indices sit in the bytes of one register (C <= 255):
// a predicated PRMT replaces the select.
template <int K> __device__ __forceinline__ void argmax_step(float& best, unsigned& bi4, float z, unsigned c, float one) {
  // The sweep sits on the ALU pipe (FSETP / FSEL / PRMT run at half rate) while the FMA pipe is a quarter busy: the value
  // update is issued as a predicated FFMA, best = z * one + (-0) == z for every z, with `one` a run-time 1.0f the
  // assembler cannot fold — it runs on the FMA pipe.
  asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %2, %0;\n\t@p fma.rn.f32 %0, %2, %5, 0f80000000;\n\t@p prmt.b32 %1, %1, %3, %4;\n\t}"
      : "+f"(best), "+r"(bi4)
      : "f"(z), "r"(c), "n"(K == 0 ? 0x3214 : (K == 1 ? 0x3240 : (K == 2 ? 0x3410 : 0x4210))), "f"(one));
}

constexpr int kPrefetchAhead = 4;
// One chunk of PXC columns (npx <= PXC of them live) of a band unit: class sweep + write-out of nrow rows.
template <typename T, int THREADS, bool PRIVATE, int R, int PXC>
__device__ __forceinline__ void band_chunk(const b200seg_image& im, const T* base, int hw, int C, int o00, int o01, int o10, int o11,
                                           const float (&h0)[R], const float (&h1)[R], float sw, bool ac, int x0, int Xc, int npx,
                                           int Y0, int nrow, const ClassDecoder& dgt, long long* pout,
                                           Counters<THREADS, PRIVATE>& ctr) {
  float w0[PXC], w1[PXC], best[R][PXC];
  unsigned bi4[PXC];
  const float one = __fmul_rn((float)blockDim.x, 1.f / THREADS);   // 1.0f, opaque to the assembler (see argmax_step)
#pragma unroll
  for (int j = 0; j < PXC; ++j) {
    const float s = aten_src_index(sw, Xc + min(j, npx - 1), ac);
    w1[j] = __fsub_rn(s, (float)x0);                       // ATen: lambda1 = src - (int)src, and (int)src == x0 in this run
    w0[j] = __fsub_rn(1.f, w1[j]);
    bi4[j] = 0u;
#pragma unroll
    for (int k = 0; k < R; ++k) best[k][j] = neg_inf();
  }
  const T* pl = base;
  T ta = __ldg(pl + o00), tb = __ldg(pl + o01), tc = __ldg(pl + o10), td = __ldg(pl + o11);
  if constexpr (PXC <= 4) {
    // the ground truth of the chunk's rows is requested into L2 now and read after the class sweep (narrow chunks only:
    // their sweep is short of covering the write-out's DRAM latency; the 8-wide form measured slower with it)
    const unsigned gb = (unsigned)label_bytes(dgt.dt);
    const char* gp = reinterpret_cast<const char*>(im.gt) + ((size_t)Y0 * im.W + Xc) * gb;
#pragma unroll
    for (int k = 0; k < R; ++k)
      if (k < nrow) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + (size_t)k * im.W * gb));
  }
  for (int c = 0; c < C; ++c) {
    const float a = to_float<T>(ta), bb = to_float<T>(tb), cc = to_float<T>(tc), d = to_float<T>(td);
    pl += hw;
    if (c + 1 < C) {                                       // the next class's taps are in flight during this class's arithmetic
      ta = __ldg(pl + o00); tb = __ldg(pl + o01); tc = __ldg(pl + o10); td = __ldg(pl + o11);
    }
    if constexpr (PXC <= 4) {
      // ... and the lines of the class kPrefetchAhead further on are requested into L1 (no register cost): at 1/4
      // resolution the logits of an image are read about once, so a tap load is a DRAM access, which one class of
      // arithmetic (~500 cycles of a scheduler's four warps) does not cover. A warp's lanes hold neighbouring runs: the
      // left taps of all lanes cover the right taps' sectors but one. (Not in the 8-wide sweep: at 1/8 resolution the
      // logits are small, the sweep is issue bound and the two extra instructions per class cost 6 %.)
      const T* pf = base + (size_t)min(c + kPrefetchAhead, C - 1) * hw;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + o00));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + o10));
    }
    float Xv[PXC], Yv[PXC];
#pragma unroll
    for (int j = 0; j < PXC; ++j) {
      Xv[j] = __fmaf_rn(w0[j], a, __fmul_rn(w1[j], bb));
      Yv[j] = __fmaf_rn(w0[j], cc, __fmul_rn(w1[j], d));
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {                          // row outside, column inside: consecutive updates hit different registers
#pragma unroll
      for (int j = 0; j < PXC; ++j) {
        float z = __fmaf_rn(h0[k], Xv[j], __fmul_rn(h1[k], Yv[j]));
        if constexpr (sizeof(T) == 2) z = to_float<T>(from_float<T>(z));   // F.interpolate returns the logit dtype
        if (k == 0) argmax_step<0>(best[k][j], bi4[j], z, (unsigned)c, one);   // strict >: lowest index wins ties
        else if (k == 1) argmax_step<1>(best[k][j], bi4[j], z, (unsigned)c, one);
        else if (k == 2) argmax_step<2>(best[k][j], bi4[j], z, (unsigned)c, one);
        else argmax_step<3>(best[k][j], bi4[j], z, (unsigned)c, one);
      }
    }
  }
  // write-out, one row at a time (a rolled loop: unrolled it was 4000 instructions of straight-line code); the next row's
  // ground truth is requested before this row's counters are updated
  int gnext[PXC];
  dgt.template upto<PXC>(im.gt, (size_t)Y0 * im.W + Xc, npx, gnext);
#pragma unroll 1
  for (int k = 0; k < nrow; ++k) {
    const size_t px0 = (size_t)(Y0 + k) * im.W + Xc;
    int gv[PXC];
#pragma unroll
    for (int j = 0; j < PXC; ++j) gv[j] = gnext[j];
    if (k + 1 < nrow) dgt.template upto<PXC>(im.gt, px0 + im.W, npx, gnext);
    const int sh8 = 8 * k;
    if (pout) {
#pragma unroll
      for (int j = 0; j < PXC; ++j)
        if (j < npx) pout[px0 + j] = (long long)((bi4[j] >> sh8) & 0xffu);
    }
#pragma unroll
    for (int j = 0; j < PXC; ++j) {
      const int bi = (int)((bi4[j] >> sh8) & 0xffu);
      if constexpr (PRIVATE) ctr.update_private(ctr.cnt + threadIdx.x, bi, gv[j]);
      else if (gv[j] != kIgnored) ctr.update(bi, gv[j]);
    }
  }
}

template <typename T, int THREADS, bool PRIVATE, int R>
__device__ __forceinline__ void resize_band_units(const b200seg_image& im, const int* run_x, const int* run_y, int G, unsigned u_begin,
                                                  unsigned u_end, float sh, float sw, bool ac, int C, const ClassDecoder& dgt,
                                                  long long* pout, Counters<THREADS, PRIVATE>& ctr) {
  static_assert(R == 4, "the row slots are the four bytes of a register");
  const unsigned runs = (unsigned)im.w + 1u, GR = (unsigned)G * runs;
  const int hw = im.h * im.w;
  const T* base = reinterpret_cast<const T*>(im.pred);
  for (unsigned uw = u_begin + (threadIdx.x & ~31u); uw < u_end; uw += THREADS) {     // warp-uniform trip count
    const unsigned u = uw + (threadIdx.x & 31u);
    const unsigned uc = u < u_end ? u : u_end - 1u;
    const unsigned b = uc / GR, rem = uc - b * GR, sgrp = rem / runs, r = rem - sgrp * runs;
    const int Y0 = run_y[b] + (int)sgrp * R, Yend = run_y[b + 1];
    const int X0 = run_x[r];
    const int X1 = (u < u_end && Y0 < Yend) ? run_x[r + 1] : X0;     // an empty unit is a run without columns
    const int nrow = max(1, min(R, Yend - Y0));
    int y0, y1, x0, x1;
    float h0[R], h1[R], lxf;
#pragma unroll
    for (int k = 0; k < R; ++k) {                          // (y0, y1) is the same for every row of the band
      resize_src(sh, Y0 + min(k, nrow - 1), im.h, ac, y0, y1, h1[k]);
      h0[k] = __fsub_rn(1.f, h1[k]);
    }
    resize_src(sw, X0 < im.W ? X0 : im.W - 1, im.w, ac, x0, x1, lxf);
    const int o00 = y0 * im.w + x0, o01 = y0 * im.w + x1, o10 = y1 * im.w + x0, o11 = y1 * im.w + x1;
    // Chunks of 8 columns; what is left of the runs takes a narrower chunk — a run of 9 columns in one lane would otherwise
    // cost its whole warp a second 8-wide class sweep with a single live column. The width is chosen per WARP (the
    // widest remainder among its lanes), so that short edge runs ride masked in their neighbours' sweep.
    int Xc = X0;
    while (true) {
      const int left = X1 - Xc;
      const int widest = __reduce_max_sync(0xffffffffu, left);
      if (widest <= 0) break;
      if (widest > 4) {
        if (left > 0)
          band_chunk<T, THREADS, PRIVATE, R, 8>(im, base, hw, C, o00, o01, o10, o11, h0, h1, sw, ac, x0, Xc, min(left, 8), Y0, nrow, dgt, pout, ctr);
        Xc += 8;
      } else if (widest > 2) {
        if (left > 0)
          band_chunk<T, THREADS, PRIVATE, R, 4>(im, base, hw, C, o00, o01, o10, o11, h0, h1, sw, ac, x0, Xc, min(left, 4), Y0, nrow, dgt, pout, ctr);
        Xc += 4;
      } else {
        if (left > 0)
          band_chunk<T, THREADS, PRIVATE, R, 2>(im, base, hw, C, o00, o01, o10, o11, h0, h1, sw, ac, x0, Xc, min(left, 2), Y0, nrow, dgt, pout, ctr);
        Xc += 2;
      }
    }
  }
}

template <typename T, int THREADS, bool PRIVATE>
__global__ void __launch_bounds__(THREADS, 2) confusion_resize_kernel(const ConfParams p, const int align_corners) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int run_x[kRunTableMax + 2];
  __shared__ int run_y[kRunTableMax + 2];
  __shared__ int band_max;
  Counters<THREADS, PRIVATE> ctr{reinterpret_cast<unsigned int*>(smem_raw), p.C};
  ctr.zero();
  __syncthreads();
  const int C = p.C;
  const long long per = (p.total_chunks + gridDim.x - 1) / gridDim.x;
  long long chunk = (long long)blockIdx.x * per;
  const long long chunk_end = (chunk + per < p.total_chunks) ? chunk + per : p.total_chunks;
  if (chunk >= chunk_end) return;
  int img = find_image(p.chunk_prefix, p.n_images, chunk);
  long long since_flush = 0;
  ClassDecoder dgt;
  dgt.init(p.gt_dtype, C, true, p.ignore);
  const bool ac = align_corners != 0;
  constexpr int PXC = 8;
  while (chunk < chunk_end) {
    while (img + 1 < p.n_images && chunk >= p.chunk_prefix[img + 1]) ++img;
    const b200seg_image im = p.images[img];
    const long long img_chunk_end = p.chunk_prefix[img + 1] < chunk_end ? p.chunk_prefix[img + 1] : chunk_end;
    const float sh = resize_scale(im.h, im.H, ac), sw = resize_scale(im.w, im.W, ac);
    long long* pout = p.pred_out ? p.pred_out[img] : nullptr;
    // logits wider than the run table, or runs too short to share anything (low ratios, down-sampling): per-pixel form
    if (im.w + 2 > kRunTableMax || sw * p.simple_ratio > 1.f) {
      const long long px_begin = (chunk - p.chunk_prefix[img]) * kChunk;
      long long px_end = (img_chunk_end - p.chunk_prefix[img]) * kChunk;
      if (px_end > im.n_pixels) px_end = im.n_pixels;
      resize_pixels_simple<T, THREADS, PRIVATE>(im, px_begin, px_end, sh, sw, ac, C, dgt, pout, ctr);
      if (!p.totals_only) {
        ctr.flush(p.areas + (size_t)img * 3 * C);
      } else {
        since_flush += img_chunk_end - chunk;
        if (since_flush >= kMaxChunksPerFlush) { ctr.flush(p.areas); since_flush = 0; }
      }
      chunk = img_chunk_end;
      continue;
    }
    const int hw = im.h * im.w;
    const T* base = reinterpret_cast<const T*>(im.pred);
    // run starts of this image: run r = columns [run_x[r], run_x[r + 1]); bands of rows likewise in run_y
    constexpr int kBandRows = 4;
    const bool band_ok = im.h + 2 <= kRunTableMax && C <= 255;
    __syncthreads();
    if (threadIdx.x == 0) band_max = 0;
    for (int r = threadIdx.x; r <= im.w + 1; r += THREADS) run_x[r] = aten_run_start(sw, r, im.w, im.W, ac);
    if (band_ok)
      for (int r = threadIdx.x; r <= im.h + 1; r += THREADS) run_y[r] = aten_run_start(sh, r, im.h, im.H, ac);
    __syncthreads();
    if (band_ok) {
      int longest = 0;
      for (int r = threadIdx.x; r <= im.h; r += THREADS) longest = max(longest, run_y[r + 1] - run_y[r]);
      if (longest > 1) atomicMax(&band_max, longest);
      __syncthreads();
    }
    const int runs = im.w + 1;
    const long long img_chunks = p.chunk_prefix[img + 1] - p.chunk_prefix[img];
    // up-sampled rows (some band holds >= 2 rows): units of (band, group of kBandRows rows, run)
    const int G = (band_max + kBandRows - 1) / kBandRows;
    const long long band_units = (long long)(im.h + 1) * G * runs;
    if (band_ok && band_max >= 2 && band_units < (1ll << 31)) {
      const long long upc = (band_units + img_chunks - 1) / img_chunks;
      const long long u_begin = (chunk - p.chunk_prefix[img]) * upc;
      long long u_end = (img_chunk_end - p.chunk_prefix[img]) * upc;
      if (u_end > band_units) u_end = band_units;
      if (u_begin < u_end)
        resize_band_units<T, THREADS, PRIVATE, kBandRows>(im, run_x, run_y, G, (unsigned)u_begin, (unsigned)u_end, sh, sw, ac, C, dgt,
                                                          pout, ctr);
      if (!p.totals_only) {
        ctr.flush(p.areas + (size_t)img * 3 * C);
      } else {
        since_flush += img_chunk_end - chunk;
        if (since_flush >= kMaxChunksPerFlush) { ctr.flush(p.areas); since_flush = 0; }
      }
      chunk = img_chunk_end;
      continue;
    }
    // this CTA's share of the image's (row, run) units: the image's chunks split the units evenly
    const long long units = (long long)im.H * runs;
    const long long upc = (units + img_chunks - 1) / img_chunks;
    const long long u_begin = (chunk - p.chunk_prefix[img]) * upc;
    long long u_end = (img_chunk_end - p.chunk_prefix[img]) * upc;
    if (u_end > units) u_end = units;
    for (long long u = u_begin + threadIdx.x; u < u_end; u += THREADS) {
      const int Y = (int)(u / runs), r = (int)(u - (long long)Y * runs);
      const int X0 = run_x[r], X1 = run_x[r + 1];
      if (X0 >= X1) continue;
      int y0, y1, x0, x1;
      float ly, lxf;
      resize_src(sh, Y, im.h, ac, y0, y1, ly);
      resize_src(sw, X0, im.w, ac, x0, x1, lxf);           // (x0, x1) is the same for every column of the run
      const float h1 = ly, h0 = __fsub_rn(1.f, ly);
      const int o00 = y0 * im.w + x0, o01 = y0 * im.w + x1, o10 = y1 * im.w + x0, o11 = y1 * im.w + x1;
      const long long row_px = (long long)Y * im.W;
      for (int Xc = X0; Xc < X1; Xc += PXC) {
        const int npx = min(PXC, X1 - Xc);
        float w0[PXC], w1[PXC], best[PXC];
        int bi[PXC];
#pragma unroll
        for (int j = 0; j < PXC; ++j) {
          const float s = aten_src_index(sw, Xc + min(j, npx - 1), ac);
          w1[j] = __fsub_rn(s, (float)x0);                 // ATen: lambda1 = src - (int)src, and (int)src == x0 in this run
          w0[j] = __fsub_rn(1.f, w1[j]);
          best[j] = neg_inf();
          bi[j] = 0;
        }
        const T* pl = base;
        for (int c = 0; c < C; ++c) {
          const float a = to_float<T>(__ldg(pl + o00)), b = to_float<T>(__ldg(pl + o01));
          const float cc = to_float<T>(__ldg(pl + o10)), d = to_float<T>(__ldg(pl + o11));
          pl += hw;
#pragma unroll
          for (int j = 0; j < PXC; ++j) {
            float z = aten_bilerp(h0, h1, w0[j], w1[j], a, b, cc, d);
            if constexpr (sizeof(T) == 2) z = to_float<T>(from_float<T>(z));   // F.interpolate returns the logit dtype
            if (z > best[j]) { best[j] = z; bi[j] = c; }                      // lowest index wins ties
          }
        }
#pragma unroll
        for (int j = 0; j < PXC; ++j) {
          const bool ok = j < npx;
          const long long px = row_px + Xc + min(j, npx - 1);
          const int gv = ok ? dgt.one(im.gt, (size_t)px) : kIgnored;
          if (ok && pout) pout[px] = bi[j];
          if constexpr (PRIVATE) ctr.update_private(ctr.cnt + threadIdx.x, bi[j], gv);
          else if (gv != kIgnored) ctr.update(bi[j], gv);
        }
      }
    }
    if (!p.totals_only) {
      ctr.flush(p.areas + (size_t)img * 3 * C);
    } else {
      since_flush += img_chunk_end - chunk;
      if (since_flush >= kMaxChunksPerFlush) { ctr.flush(p.areas); since_flush = 0; }
    }
    chunk = img_chunk_end;
  }
  if (p.totals_only) ctr.flush(p.areas);
}

// Persistent grid: exactly (SMs x resident CTAs per SM), so the static chunk partition has no tail wave.
template <typename K> static int persistent_grid(K kernel, int threads, size_t smem, long long total_chunks, int* grid) {
  int per_sm = 0;
  B200SEG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  if (per_sm < 1) per_sm = 1;
  long long g = (long long)kSMs * per_sm;
  if (g > total_chunks) g = total_chunks;
  *grid = (int)(g < 1 ? 1 : g);
  return 0;
}

template <typename T, bool FROM_LOGITS> static int launch_confusion(const ConfParams& p, cudaStream_t st) {
  // private per-thread counters when they fit in shared memory, else shared atomics
  const size_t need256 = (size_t)3 * p.C * 256 * 4, need128 = (size_t)3 * p.C * 128 * 4;
  int grid = 1;
  if (need256 <= 72 * 1024) {
    auto k = confusion_kernel<T, 256, true, FROM_LOGITS>;
    if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(k), 72 * 1024)) return e;
    if (int e = persistent_grid(k, 256, need256, p.total_chunks, &grid)) return e;
    k<<<grid, 256, need256, st>>>(p);
  } else if (need128 <= 200 * 1024) {
    auto k = confusion_kernel<T, 128, true, FROM_LOGITS>;
    if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(k), 200 * 1024)) return e;
    if (int e = persistent_grid(k, 128, need128, p.total_chunks, &grid)) return e;
    k<<<grid, 128, need128, st>>>(p);
  } else {
    auto k = confusion_kernel<T, 256, false, FROM_LOGITS>;
    if (int e = persistent_grid(k, 256, (size_t)3 * p.C * 4, p.total_chunks, &grid)) return e;
    k<<<grid, 256, (size_t)3 * p.C * 4, st>>>(p);
  }
  count_launch();
  return check_launch("confusion_kernel");
}

template <typename T> static int launch_confusion_resize(ConfParams p, int align_corners, cudaStream_t st) {
  {
    const char* e = getenv("B200SEG_F2_SIMPLE_RATIO");   // A/B measurements
    p.simple_ratio = e ? (float)atof(e) : kSimpleRatio;
  }
  const size_t need256 = (size_t)3 * p.C * 256 * 4;
  int grid = 1;
  if (need256 <= 72 * 1024) {
    auto k = confusion_resize_kernel<T, 256, true>;
    if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(k), 72 * 1024)) return e;
    if (int e = persistent_grid(k, 256, need256, p.total_chunks, &grid)) return e;
    k<<<grid, 256, need256, st>>>(p, align_corners);
  } else {
    auto k = confusion_resize_kernel<T, 256, false>;
    if (int e = persistent_grid(k, 256, (size_t)3 * p.C * 4, p.total_chunks, &grid)) return e;
    k<<<grid, 256, (size_t)3 * p.C * 4, st>>>(p, align_corners);
  }
  count_launch();
  return check_launch("confusion_resize_kernel");
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int32_t b200seg_confusion_chunk_pixels(void) { return kChunk; }

extern "C" int b200seg_confusion_logits_resized(const b200seg_image* images, const int64_t* chunk_prefix, int32_t n_images,
                                                int64_t total_chunks, int32_t chunk_pixels, int32_t logit_dtype,
                                                int32_t gt_dtype, int32_t C, int64_t ignore_index, int32_t align_corners,
                                                int64_t* areas, int64_t* const* pred_out, int32_t totals_only, void* stream) {
  B200SEG_REQUIRE(chunk_pixels == kChunk, "confusion: chunk_pixels must be %d", kChunk);
  B200SEG_REQUIRE(C >= 1 && C <= 4096, "confusion: num_classes %d out of range [1,4096]", C);
  B200SEG_REQUIRE(n_images >= 0 && areas, "confusion: bad arguments");
  if (n_images == 0 || total_chunks == 0) return 0;
  B200SEG_REQUIRE(images && chunk_prefix, "confusion: NULL image table");
  ConfParams p{images, (const long long*)chunk_prefix, n_images, total_chunks, 0, gt_dtype, C,
               ignore_index, (long long*)areas, (long long* const*)pred_out, totals_only};
  cudaStream_t st = (cudaStream_t)stream;
  switch (logit_dtype) {
    case B200SEG_F32: return launch_confusion_resize<float>(p, align_corners, st);
    case B200SEG_BF16: return launch_confusion_resize<__nv_bfloat16>(p, align_corners, st);
    case B200SEG_F16: return launch_confusion_resize<__half>(p, align_corners, st);
  }
  set_error("confusion: unsupported logit dtype %d", logit_dtype);
  return 1;
}

extern "C" int b200seg_confusion_labels(const b200seg_image* images, const int64_t* chunk_prefix, int32_t n_images,
                                        int64_t total_chunks, int32_t chunk_pixels, int32_t pred_dtype,
                                        int32_t gt_dtype, int32_t C, int64_t ignore_index, int64_t* areas,
                                        int32_t totals_only, void* stream) {
  B200SEG_REQUIRE(chunk_pixels == kChunk, "confusion: chunk_pixels must be %d", kChunk);
  B200SEG_REQUIRE(C >= 1 && C <= 4096, "confusion: num_classes %d out of range [1,4096]", C);
  B200SEG_REQUIRE(n_images >= 0 && areas, "confusion: bad arguments");
  if (n_images == 0 || total_chunks == 0) return 0;
  B200SEG_REQUIRE(images && chunk_prefix, "confusion: NULL image table");
  ConfParams p{images, (const long long*)chunk_prefix, n_images, total_chunks, pred_dtype, gt_dtype, C,
               ignore_index, (long long*)areas, nullptr, totals_only};
  return launch_confusion<float, false>(p, (cudaStream_t)stream);
}

extern "C" int b200seg_confusion_logits(const b200seg_image* images, const int64_t* chunk_prefix, int32_t n_images,
                                        int64_t total_chunks, int32_t chunk_pixels, int32_t logit_dtype,
                                        int32_t gt_dtype, int32_t C, int64_t ignore_index, int64_t* areas,
                                        int64_t* const* pred_out, int32_t totals_only, void* stream) {
  B200SEG_REQUIRE(chunk_pixels == kChunk, "confusion: chunk_pixels must be %d", kChunk);
  B200SEG_REQUIRE(C >= 1 && C <= 4096, "confusion: num_classes %d out of range [1,4096]", C);
  B200SEG_REQUIRE(n_images >= 0 && areas, "confusion: bad arguments");
  if (n_images == 0 || total_chunks == 0) return 0;
  B200SEG_REQUIRE(images && chunk_prefix, "confusion: NULL image table");
  ConfParams p{images, (const long long*)chunk_prefix, n_images, total_chunks, 0, gt_dtype, C,
               ignore_index, (long long*)areas, (long long* const*)pred_out, totals_only};
  cudaStream_t st = (cudaStream_t)stream;
  switch (logit_dtype) {
    case B200SEG_F32: return launch_confusion<float, true>(p, st);
    case B200SEG_BF16: return launch_confusion<__nv_bfloat16, true>(p, st);
    case B200SEG_F16: return launch_confusion<__half, true>(p, st);
  }
  set_error("confusion: unsupported logit dtype %d", logit_dtype);
  return 1;
}
