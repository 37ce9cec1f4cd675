// Class-split REGISTER-TILE kernels for logits at label resolution, sm_100a:
//   rt_fwd_kernel<MODE_GRAD>   cross-entropy (+accuracy) forward AND backward in one pass  (b200seg_loss_fused_fwdbwd)
//   rt_fwd_kernel<MODE_DICE>   Dice partial sums + cross-entropy + accuracy forward        (b200seg_loss_fwd, WANT_DICE)
//   rt_dice_bwd_kernel         Dice + cross-entropy backward                               (b200seg_loss_bwd,  WANT_DICE)
//
// Replaces (reference file:line) DiceLoss.forward models/losses/dice_loss.py:103-134 — F.softmax, F.one_hot (an
// int64 (N,H,W,C) tensor: 5 GB at ADE20K shape) and the per-class Python loop dice_loss/binary_dice_loss :23-58
// (>= 7 launches per class) — together with F.cross_entropy / accuracy (cross_entropy_loss.py:56-61,
// accuracy.py:41-60) and their autograd backwards, with ONE read of the logits per direction.
//
// Both Dice (sum_px softmax(z)_c^e for every class) and the single-pass CE gradient need the NORMALISED probability
// of every (pixel, class) element, so the exponentials must be kept until the per-pixel sum is known. A thread keeps
// CPT >= C classes x V pixels in registers: one warp owns all classes of its 32*V pixels, so there is no inter-warp
// traffic at all (4 independent warps per CTA). Per-class Dice sums live in per-thread registers across all tiles a
// CTA visits: one shuffle tree per class per CTA lifetime. One MUFU.EX2 per element. More than 32 classes take the
// streaming kernels (loss_dice.cu); a class-split variant of this file (warps exchanging partial max / sum through
// shared memory) was measured at 9 % of the roofline and removed.
//
// Roofline: HBM. Algorithmic bytes: forward el*s + px*L (+4 px for lse); backward / single pass 2*el*s + px*L.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace b200seg {

constexpr float kPad = -1.0e30f;
enum { MODE_GRAD = 0, MODE_DICE = 1 };

}  // namespace b200seg
#include "loss_rt_params.cuh"
namespace b200seg {

// 4 resident CTAs per SM (register cap 128): measured 0.59 / 0.65 / 0.73 of the roofline at 2 / 3 / 4 CTAs, 0.54 at 5 (spills)
template <typename T, int V, int CPT, int MODE>
__global__ void __launch_bounds__(128, 4) rt_fwd_kernel(const RtParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  static_assert(CPT <= 32, "one lane per class in the flush");
  constexpr int PXW = 32 * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = p.C;
  constexpr int NPG = 4;                   // pixel groups (warps) per CTA
  const int pgw = warp;
  constexpr int c0 = 0;
  const int c1 = C;
  const int n = blockIdx.y;
  const long long HW = p.HW;

  // shared memory: per-warp one-hot Dice bins
  float* A_s = reinterpret_cast<float*>(smem_raw);   // [warps][C] (MODE_DICE)
  const int nwarps = blockDim.x >> 5;
  float* T_s = A_s + (MODE == MODE_DICE ? nwarps * C : 0);
  if constexpr (MODE == MODE_DICE) {
    for (int i = threadIdx.x; i < 2 * nwarps * C; i += blockDim.x) A_s[i] = 0.f;
    __syncthreads();
  }

  const T* img = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * HW;
  T* gimg = reinterpret_cast<T*>(p.grad) + (size_t)n * C * HW;
  const bool want_ce = (p.flags & B200SEG_WANT_CE) != 0;
  const bool e2 = (p.dice_exponent == 2.f);
  float Gs = 0.f;
  if constexpr (MODE == MODE_GRAD) Gs = p.ce_scale_host * (p.ce_grad_out ? __ldg(p.ce_grad_out) : 1.f);

  float accB[MODE == MODE_DICE ? CPT : 1];
#pragma unroll
  for (int i = 0; i < (MODE == MODE_DICE ? CPT : 1); ++i) accB[i] = 0.f;
  float loss_acc = 0.f;
  int n_valid = 0, n_correct = 0, n_bad = 0, n_acc = 0;

  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const long long px0 = ((long long)tile * NPG + pgw) * PXW + (long long)lane * V;
    const bool active = px0 < HW;
    constexpr bool owner = true;
    float z[CPT][V];
    {
      // running pointer over the class dimension: one 64-bit add per class instead of a 64-bit multiply-add. The host
      // picks the smallest CPT >= C, so classes below kSure always exist and carry no predicate; lanes past the end of
      // the image read the tile's first pixel again (their results are masked by `active`).
      constexpr int kSure = V == 2 ? (CPT == 32 ? 24 : (CPT >= 8 ? CPT - 4 : 0)) : (CPT == 32 ? 8 : 0);
      const T* q = img + (size_t)c0 * HW + (active ? px0 : 0);
      const int ncls = c1 - c0;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        if (i < kSure || i < ncls) {
          load_vec<T, V>(q, z[i]);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) z[i][v] = kPad;
        }
        q += HW;
      }
    }
    long long y[V];
    float pwv[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { y[v] = p.ignore_index; pwv[v] = 1.f; }
    if (owner && active) {
      load_labels<V>(p.labels, p.label_dtype, (size_t)n * HW + px0, y);
      if (p.pw && want_ce) load_vec<float, V>(p.pw + (size_t)n * HW + px0, pwv);
    }
    // ---- local soft-max over this thread's classes
    float m[V], s[V];
    int idx[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float lm = z[0][v];
#pragma unroll
      for (int i = 1; i < CPT; ++i) lm = fmaxf(lm, z[i][v]);           // FMNMX3 tree
      int li = c0 + CPT - 1;
#pragma unroll
      for (int i = CPT - 2; i >= 0; --i) li = (z[i][v] == lm) ? c0 + i : li;   // lowest index among the maxima
      const float nm = -lm * kLog2e;
      float ls = 0.f;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        z[i][v] = ex2(fmaf(z[i][v], kLog2e, nm));   // padded classes: 2^(-huge) = 0
        ls += z[i][v];
      }
      m[v] = lm; s[v] = ls; idx[v] = li;
    }
    // per-pixel CE coefficient and label of this tile (owner), shared with the other class groups in MODE_GRAD
    float kk[V];
    int yc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const long long yy = y[v];
      const bool valid = owner && active && (yy != p.ignore_index) && yy >= 0 && yy < (long long)C;
      kk[v] = valid ? pwv[v] * (p.cw ? __ldg(p.cw + yy) : 1.f) : 0.f;
      yc[v] = valid ? (int)yy : -1;
    }
    float f[V];
#pragma unroll
    for (int v = 0; v < V; ++v) f[v] = fast_rcp(s[v]);

    if (active) {
      if constexpr (MODE == MODE_DICE) {
        if (e2) {
#pragma unroll
          for (int i = 0; i < CPT; ++i) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float pr = z[i][v] * f[v];
              accB[i] = fmaf(pr, pr, accB[i]);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < CPT; ++i) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float pr = z[i][v] * f[v];
              accB[i] += pr > 0.f ? __powf(pr, p.dice_exponent) : 0.f;
            }
          }
        }
      } else {
        float kg[V], rr[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { kg[v] = kk[v] * Gs; rr[v] = kg[v] * f[v]; }
        T* gq = gimg + (size_t)c0 * HW + px0;
        const int ncls = c1 - c0;
        constexpr int kSureG = V == 2 ? (CPT == 32 ? 24 : (CPT >= 8 ? CPT - 4 : 0)) : (CPT == 32 ? 8 : 0);
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          if (i < kSureG || i < ncls) {
            float gr[V];
#pragma unroll
            for (int v = 0; v < V; ++v) gr[v] = rr[v] * z[i][v];
            store_vec<T, V>(gq, gr);
          }
          gq += HW;
        }
        // one-hot term: instead of a compare per element, the label's class is re-stored by the same thread (program
        // order) with k * (p_y - 1); p_y is recomputed from the label's logit with the operations of the class loop
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const int yl = yc[v] - c0;
          if (yc[v] >= 0 && yl >= 0 && yl < ncls) {
            const float zy = to_float<T>(img[(size_t)yc[v] * HW + px0 + v]);
            const float ey = ex2(fmaf(zy, kLog2e, -m[v] * kLog2e));
            gimg[(size_t)yc[v] * HW + px0 + v] = from_float<T>(rr[v] * ey - kg[v]);
          }
        }
      }
    }

    if (owner) {  // per-pixel terms: CE, accuracy, one-hot Dice sums
      float lse[V], lpx[V], pyv[V];
      int ycl[V];
#pragma unroll
      for (int v = 0; v < V; ++v) { lse[v] = 0.f; lpx[v] = 0.f; pyv[v] = 0.f; ycl[v] = -1; }
      if (active) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          lse[v] = m[v] + fast_log(s[v]);
          const long long yy = y[v];
          const bool ign = (yy == p.ignore_index);
          const bool inr = (yy >= 0 && yy < (long long)C);
          const int ycc = yy < 0 ? 0 : (yy >= (long long)C ? C - 1 : (int)yy);   // torch.clamp, dice_loss.py:120
          float zy = 0.f;
          if (MODE == MODE_DICE || (want_ce && !ign && inr)) zy = to_float<T>(img[(size_t)ycc * HW + px0 + v]);
          if (want_ce) {
            n_bad += (!ign && !inr);
            n_valid += !ign;
            const float l = kk[v] * (lse[v] - zy);   // kk = 0 unless the label is valid
            lpx[v] = l * p.lw;
            loss_acc += l;
          }
          if constexpr (MODE == MODE_DICE) {
            ycl[v] = ycc;
            pyv[v] = (yy != p.dice_ignore) ? ex2((zy - lse[v]) * kLog2e) : 0.f;   // valid_mask, dice_loss.py:122
          }
          const bool av = p.acc_has_ignore ? (yy != p.acc_ignore) : true;
          n_acc += av;
          n_correct += (av && (long long)idx[v] == yy);
        }
        if (p.lse_out) store_vec<float, V>(p.lse_out + (size_t)n * HW + px0, lse);
        if (p.loss_px) store_vec<float, V>(p.loss_px + (size_t)n * HW + px0, lpx);
      }
      if constexpr (MODE == MODE_DICE) {
        // warp-aggregated scatter of (p_y * valid, 1) into this warp's class bins
#pragma unroll
        for (int v = 0; v < V; ++v) {
          onehot_bins_add(A_s + warp * C, T_s + warp * C, ycl[v], pyv[v], lane);
        }
      }
    }
  }

  if constexpr (MODE == MODE_DICE) {
    float mineB = 0.f;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const float tot = warp_sum(accB[i]);
      if (lane == i) mineB = tot;
    }
    if (lane < CPT && c0 + lane < c1) atomicAdd(p.dice_part + ((size_t)n * C + c0 + lane) * 3 + 1, (double)mineB);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a = 0.f, t = 0.f;
      for (int q = 0; q < nwarps; ++q) { a += A_s[q * C + c]; t += T_s[q * C + c]; }
      if (t != 0.f) {
        atomicAdd(p.dice_part + ((size_t)n * C + c) * 3 + 0, (double)a);
        atomicAdd(p.dice_part + ((size_t)n * C + c) * 3 + 2, (double)t);
      }
    }
  }
  cta_flush_stats(loss_acc, n_valid, n_correct, n_bad, n_acc, p.stats, want_ce || MODE == MODE_GRAD);
}

// ------------------------------------------------------------------------------------------------
// Backward: grad_z_j = p_j * (g_j - sum_c p_c g_c) [dice, g = dL/dp]  +  k * (p_j - onehot_j) [CE]
template <typename T, int V, int CPT>
__global__ void __launch_bounds__(128) rt_dice_bwd_kernel(const RtParams p) {
  constexpr int PXW = 32 * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = p.C;
  constexpr int NPG = 4;
  const int pgw = warp;
  constexpr int c0 = 0;
  const int c1 = C;
  const int n = blockIdx.y;
  const long long HW = p.HW;

  const bool want_ce = (p.flags & B200SEG_WANT_CE) != 0;
  const bool e2 = (p.dice_exponent == 2.f);
  float Gce = 0.f;
  if (want_ce) {
    Gce = p.ce_scale_host;
    if (p.ce_grad_out) Gce *= __ldg(p.ce_grad_out);
    if (p.ce_use_nvalid) {
      const double nv = (double)(long long)p.stats[B200SEG_ST_N_VALID];
      Gce = (float)((double)Gce / (nv + 1.1920928955078125e-07));
    }
  }
  const float god = p.dice_grad_out ? __ldg(p.dice_grad_out) : 1.f;
  float beta[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i)
    beta[i] = (c0 + i < c1) ? p.dice_exponent * god * __ldg(p.dice_coef + ((size_t)n * C + c0 + i) * 2 + 1) : 0.f;

  const T* img = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * HW;
  T* gimg = reinterpret_cast<T*>(p.grad) + (size_t)n * C * HW;

  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const long long px0 = ((long long)tile * NPG + pgw) * PXW + (long long)lane * V;
    const bool active = px0 < HW;
    constexpr bool owner = true;
    float z[CPT][V];
    {
      // running pointer over the class dimension: one 64-bit add per class instead of a 64-bit multiply-add
      const T* q = img + (size_t)c0 * HW + px0;
      const int ncls = active ? c1 - c0 : 0;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        if (i < ncls) {
          load_vec<T, V>(q, z[i]);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) z[i][v] = kPad;
        }
        q += HW;
      }
    }
    float nl[V];
    {
      float lse[V];
#pragma unroll
      for (int v = 0; v < V; ++v) lse[v] = 0.f;
      if (active) load_vec<float, V>(p.lse_in + (size_t)n * HW + px0, lse);
#pragma unroll
      for (int v = 0; v < V; ++v) nl[v] = -lse[v] * kLog2e;
    }
    float dotp[V];
#pragma unroll
    for (int v = 0; v < V; ++v) dotp[v] = 0.f;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float pr = ex2(fmaf(z[i][v], kLog2e, nl[v]));
        z[i][v] = pr;
        if (e2) {
          dotp[v] = fmaf(beta[i] * pr, pr, dotp[v]);
        } else {
          const float gd = pr > 0.f ? beta[i] * __powf(pr, p.dice_exponent - 1.f) : 0.f;
          dotp[v] = fmaf(gd, pr, dotp[v]);
        }
      }
    }
    float kk[V], da[V];
    int yc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { kk[v] = 0.f; da[v] = 0.f; yc[v] = -1; }
    if (owner && active) {
      long long y[V];
      float pwv[V], gpx[V];
#pragma unroll
      for (int v = 0; v < V; ++v) { pwv[v] = 1.f; gpx[v] = 1.f; }
      load_labels<V>(p.labels, p.label_dtype, (size_t)n * HW + px0, y);
      if (want_ce && p.pw) load_vec<float, V>(p.pw + (size_t)n * HW + px0, pwv);
      if (want_ce && p.ce_grad_px) load_vec<float, V>(p.ce_grad_px + (size_t)n * HW + px0, gpx);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const long long yy = y[v];
        yc[v] = yy < 0 ? 0 : (yy >= (long long)C ? C - 1 : (int)yy);
        if (yy != p.dice_ignore) {
          da[v] = god * __ldg(p.dice_coef + ((size_t)n * C + yc[v]) * 2 + 0);
          const float zy = to_float<T>(img[(size_t)yc[v] * HW + px0 + v]);
          dotp[v] -= da[v] * ex2(fmaf(zy, kLog2e, nl[v]));
        }
        const bool valid = (yy != p.ignore_index) && yy >= 0 && yy < (long long)C;
        if (want_ce && valid) kk[v] = Gce * pwv[v] * gpx[v] * (p.cw ? __ldg(p.cw + yy) : 1.f);
      }
    }
    float sub[V];
#pragma unroll
    for (int v = 0; v < V; ++v) sub[v] = kk[v] - dotp[v];   // grad = p * (gd + k - dot)
    if (active) {
      T* gq = gimg + (size_t)c0 * HW + px0;
      const int ncls = c1 - c0;
      int yl[V];
#pragma unroll
      for (int v = 0; v < V; ++v) yl[v] = yc[v] - c0;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        if (i < ncls) {
          float gr[V];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float pr = z[i][v];
            float gd;
            if (e2) gd = beta[i] * pr;
            else gd = pr > 0.f ? beta[i] * __powf(pr, p.dice_exponent - 1.f) : 0.f;
            float gv = pr * (gd + sub[v]);
            if (i == yl[v]) gv -= fmaf(pr, da[v], kk[v]);
            gr[v] = gv;
          }
          store_vec<T, V>(gq, gr);
        }
        gq += HW;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int resident_ctas(int threads, int regs_hint) {
  const int by_regs = 65536 / (threads * regs_hint);
  const int by_threads = 2048 / threads;
  int r = by_regs < by_threads ? by_regs : by_threads;
  return r < 1 ? 1 : r;
}

template <typename T, int V, int CPT, int MODE> static int launch_rt_fwd(RtParams p, cudaStream_t st) {
  B200SEG_REQUIRE(p.C <= CPT, "register-tile kernels hold at most %d classes per warp (got %d)", CPT, p.C);
  const long long per_tile = 4LL * 32 * V;   // 4 warps x 32 lanes x V pixels
  p.tiles = (int)((p.HW + per_tile - 1) / per_tile);
  const int threads = 128;
  const size_t smem = MODE == MODE_DICE ? (size_t)2 * (threads / 32) * p.C * 4 : 0;
  int gx = (kSMs * resident_ctas(threads, 128) * 2 + p.N - 1) / p.N;
  if (gx > p.tiles) gx = p.tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, p.N);
  rt_fwd_kernel<T, V, CPT, MODE><<<grid, threads, smem, st>>>(p);
  count_launch();
  return check_launch("rt_fwd_kernel");
}

template <typename T, int V, int CPT> static int launch_rt_bwd(RtParams p, cudaStream_t st) {
  B200SEG_REQUIRE(p.C <= CPT, "register-tile kernels hold at most %d classes per warp (got %d)", CPT, p.C);
  const long long per_tile = 4LL * 32 * V;
  p.tiles = (int)((p.HW + per_tile - 1) / per_tile);
  const int threads = 128;
  int gx = (kSMs * resident_ctas(threads, 128) * 2 + p.N - 1) / p.N;
  if (gx > p.tiles) gx = p.tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, p.N);
  rt_dice_bwd_kernel<T, V, CPT><<<grid, threads, 0, st>>>(p);
  count_launch();
  return check_launch("rt_dice_bwd_kernel");
}

// classes per thread: the smallest instantiation that holds all C <= 32 classes in one warp. The scalar (V=1) fallback
// for odd H*W only has 8 / 32.
static int pick_cpt(int C, int V) {
  if (V == 1) return C <= 8 ? 8 : 32;
  if (C <= 32) return C <= 4 ? 4 : (C <= 8 ? 8 : (C <= 12 ? 12 : (C <= 16 ? 16 : (C <= 20 ? 20 : (C <= 24 ? 24 : 32)))));
  return 32;
}

#define B200SEG_RT_CASE(CPTV)                                                                      \
  case CPTV:                                                                                       \
    if (kind == 0) return launch_rt_fwd<T, V, CPTV, MODE_GRAD>(p, st);                             \
    if (kind == 1) return launch_rt_fwd<T, V, CPTV, MODE_DICE>(p, st);                             \
    return launch_rt_bwd<T, V, CPTV>(p, st);

template <typename T, int V> static int rt_pick(const RtParams& p, int kind, cudaStream_t st) {
  switch (pick_cpt(p.C, V)) {
    B200SEG_RT_CASE(8)
    B200SEG_RT_CASE(32)
    default: break;
  }
  if constexpr (V == 2) {
    switch (pick_cpt(p.C, V)) {
      B200SEG_RT_CASE(4)
      B200SEG_RT_CASE(12)
      B200SEG_RT_CASE(16)
      B200SEG_RT_CASE(20)
      B200SEG_RT_CASE(24)
      default: break;
    }
  }
  set_error("internal: no register-tile instantiation for C=%d", p.C);
  return 1;
}

template <typename T> int rt_run(const RtParams& p, int kind, bool vec, cudaStream_t st) {
  return vec ? rt_pick<T, 2>(p, kind, st) : rt_pick<T, 1>(p, kind, st);
}

}  // namespace b200seg
