// Dice (+ cross-entropy + accuracy) forward / backward: class-split register tiles, sm_100a.
//
// Replaces (reference file:line) DiceLoss.forward models/losses/dice_loss.py:103-134 — F.softmax,
// F.one_hot (an int64 (N,H,W,C) tensor: 5 GB at ADE20K shape), and the per-class Python loop
// dice_loss/binary_dice_loss :23-58 (>= 7 launches per class) — together with the CE / accuracy
// chain of loss_stream.cu, in ONE read of the logits.
//
// Dice needs sum_px softmax(z)_c^e for every class, i.e. the normalised probability of every
// (pixel, class) element — so unlike plain CE the exponentials must be kept until the per-pixel
// sum is known. Holding C=150 values per pixel in one thread is impossible; instead the class
// dimension is split over G warps ("class groups"): warp g of a pixel group holds classes
// [g*cpg, (g+1)*cpg) of the same 32*V pixels in registers, and the G warps exchange only their
// per-pixel partial max / partial sum through shared memory (2 named barriers per tile). Each
// thread keeps private per-class accumulators across all the tiles it visits, so the per-class
// pixel reduction costs one warp-shuffle tree per class per CTA lifetime. One MUFU.EX2 per element.
//
// Roofline: HBM (bf16 C=150 sits close to the MUFU limit). Algorithmic bytes as in loss_stream.cu.
#include <stdlib.h>

#include "common.cuh"

namespace b200seg {

struct TileParams {
  const void* logits;
  const void* labels;
  const float* pw;
  const float* cw;
  float* lse_out;
  float* loss_px;
  unsigned long long* stats;
  double* dice_part;
  // backward
  const float* lse_in;
  const float* ce_grad_out;
  const float* ce_grad_px;
  const float* dice_coef;
  const float* dice_grad_out;
  void* grad;
  float ce_scale_host;
  int ce_use_nvalid;
  //
  int label_dtype;
  int N, C;
  long long HW;
  int flags;
  long long ignore_index;
  int acc_has_ignore;
  long long acc_ignore;
  long long dice_ignore;
  float dice_exponent;
  float lw;
  int G, cpg, PG, tiles;
};

__device__ __forceinline__ void group_barrier(int pg, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(pg + 1), "r"(nthreads) : "memory");
}

template <typename T, int V, int CPT>
__global__ void __launch_bounds__(512) tile_fwd_kernel(const TileParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double sred[5 * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = p.G, PG = p.PG, C = p.C;
  const int pg = warp / G, g = warp - pg * G;
  const int c0 = g * p.cpg;
  const int c1 = (c0 + p.cpg < C) ? c0 + p.cpg : C;
  const int n = blockIdx.y;
  const long long HW = p.HW;
  constexpr int PXW = 32 * V;  // pixels per pixel-group tile

  float* exch_m = reinterpret_cast<float*>(smem_raw);          // [PG][G][PXW]
  int* exch_i = reinterpret_cast<int*>(exch_m + PG * G * PXW);   // [PG][G][PXW]
  float* exch_s = reinterpret_cast<float*>(exch_i + PG * G * PXW);
  float* A_s = exch_s + PG * G * PXW;                           // [PG][C]
  float* T_s = A_s + PG * C;                                    // [PG][C]
  for (int i = threadIdx.x; i < 2 * PG * C; i += blockDim.x) A_s[i] = 0.f;
  __syncthreads();

  const T* img = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * HW;
  const bool e2 = (p.dice_exponent == 2.f);
  const bool want_ce = (p.flags & B200SEG_WANT_CE) != 0;

  float accB[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) accB[i] = 0.f;
  float loss_acc = 0.f;
  int n_valid = 0, n_correct = 0, n_bad = 0, n_acc = 0;

  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const long long px0 = ((long long)tile * PG + pg) * PXW + (long long)lane * V;
    const bool active = px0 < HW;
    float z[CPT][V];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      if (active && c0 + i < c1) {
        load_vec<T, V>(img + (size_t)(c0 + i) * HW + px0, z[i]);
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) z[i][v] = neg_inf();
      }
    }
    // issue the label load early (only class-group 0 consumes it)
    long long y[V];
    if (g == 0 && active) load_labels<V>(p.labels, p.label_dtype, (size_t)n * HW + px0, y);

    float m[V];
    int idx[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float lm = neg_inf();
      int li = c0;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        if (z[i][v] > lm) { lm = z[i][v]; li = c0 + i; }
      }
      m[v] = lm;
      idx[v] = li;
    }
    if (G > 1) {
      const int slot = (pg * G + g) * PXW + lane * V;
#pragma unroll
      for (int v = 0; v < V; ++v) { exch_m[slot + v] = m[v]; exch_i[slot + v] = idx[v]; }
      group_barrier(pg, 32 * G);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float bm = neg_inf();
        int bi = 0;
        for (int gg = 0; gg < G; ++gg) {
          const float xm = exch_m[(pg * G + gg) * PXW + lane * V + v];
          if (xm > bm) { bm = xm; bi = exch_i[(pg * G + gg) * PXW + lane * V + v]; }
        }
        m[v] = bm;
        idx[v] = bi;
      }
    }
    float s[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float nm = active ? -m[v] * kLog2e : 0.f;
      float ls = 0.f;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        z[i][v] = ex2(fmaf(z[i][v], kLog2e, nm));  // -inf -> 0 for padded classes / inactive lanes
        ls += z[i][v];
      }
      s[v] = ls;
    }
    if (G > 1) {
      const int slot = (pg * G + g) * PXW + lane * V;
#pragma unroll
      for (int v = 0; v < V; ++v) exch_s[slot + v] = s[v];
      group_barrier(pg, 32 * G);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float t = 0.f;
        for (int gg = 0; gg < G; ++gg) t += exch_s[(pg * G + gg) * PXW + lane * V + v];
        s[v] = t;
      }
    }
    float inv[V];
#pragma unroll
    for (int v = 0; v < V; ++v) inv[v] = active ? 1.f / s[v] : 0.f;
    if (e2) {
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float pr = z[i][v] * inv[v];
          accB[i] = fmaf(pr, pr, accB[i]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float pr = z[i][v] * inv[v];
          accB[i] += pr > 0.f ? __powf(pr, p.dice_exponent) : 0.f;
        }
      }
    }

    if (g == 0) {  // per-pixel terms: CE, accuracy, one-hot dice sums
      float lse[V], lpx[V], pwv[V];
      int ycl[V];
      float pyv[V];
#pragma unroll
      for (int v = 0; v < V; ++v) { lse[v] = 0.f; lpx[v] = 0.f; pwv[v] = 1.f; ycl[v] = -1; pyv[v] = 0.f; }
      if (active) {
        if (p.pw && want_ce) {
          if constexpr (V <= 4) load_vec<float, V>(p.pw + (size_t)n * HW + px0, pwv);
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
          lse[v] = m[v] + logf(s[v]);
          const long long yy = y[v];
          const int yc = yy < 0 ? 0 : (yy >= (long long)C ? C - 1 : (int)yy);  // torch.clamp, dice_loss.py:120
          const float zy = to_float<T>(img[(size_t)yc * HW + px0 + v]);
          const float py = ex2(fmaf(zy, kLog2e, -m[v] * kLog2e)) * inv[v];
          ycl[v] = yc;
          pyv[v] = (yy != p.dice_ignore) ? py : 0.f;  // valid_mask, dice_loss.py:122
          if (want_ce) {
            const bool ign = (yy == p.ignore_index);
            const bool inr = (yy >= 0 && yy < (long long)C);
            n_bad += (!ign && !inr);
            n_valid += !ign;
            float l = 0.f;
            if (!ign && inr) {
              const float wt = p.cw ? __ldg(p.cw + yy) : 1.f;
              l = wt * (lse[v] - zy) * pwv[v];
            }
            lpx[v] = l * p.lw;
            loss_acc += l;
          }
          const bool av = p.acc_has_ignore ? (yy != p.acc_ignore) : true;
          n_acc += av;
          n_correct += (av && (long long)idx[v] == yy);
        }
        if (p.lse_out) store_vec<float, V>(p.lse_out + (size_t)n * HW + px0, lse);
        if (p.loss_px) store_vec<float, V>(p.loss_px + (size_t)n * HW + px0, lpx);
      }
      // warp-aggregated scatter of (p_y * valid, 1) into per-warp class bins
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int cls = ycl[v];
        unsigned rem = __ballot_sync(0xffffffffu, cls >= 0);
        while (rem) {
          const int leader = __ffs(rem) - 1;
          const int lc = __shfl_sync(0xffffffffu, cls, leader);
          const bool mine = (cls == lc);
          const float sum = warp_sum(mine ? pyv[v] : 0.f);
          const unsigned mm = __ballot_sync(0xffffffffu, mine);
          if (lane == leader) {
            A_s[pg * C + lc] += sum;
            T_s[pg * C + lc] += (float)__popc(mm);
          }
          rem &= ~mm;
        }
      }
    }
  }

  // ---- flush per-class sums (one shuffle tree per class per CTA lifetime)
  {
    static_assert(CPT <= 32, "one lane per class in the flush");
    float mineB = 0.f;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const float tot = warp_sum(accB[i]);
      if (lane == i) mineB = tot;
    }
    if (lane < CPT && c0 + lane < c1) atomicAdd(p.dice_part + ((size_t)n * C + c0 + lane) * 3 + 1, (double)mineB);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, t = 0.f;
    for (int q = 0; q < PG; ++q) { a += A_s[q * C + c]; t += T_s[q * C + c]; }
    if (t != 0.f) {
      atomicAdd(p.dice_part + ((size_t)n * C + c) * 3 + 0, (double)a);
      atomicAdd(p.dice_part + ((size_t)n * C + c) * 3 + 2, (double)t);
    }
  }
  double r[5] = {(double)loss_acc, (double)n_valid, (double)n_correct, (double)n_bad, (double)n_acc};
  block_sum<double, 5>(r, sred);
  if (threadIdx.x == 0) {
    if (want_ce) {
      atomicAdd(reinterpret_cast<double*>(p.stats + B200SEG_ST_CE_SUM), r[0]);
      atomicAdd(p.stats + B200SEG_ST_N_VALID, (unsigned long long)r[1]);
      if (r[3] != 0.0) atomicAdd(p.stats + B200SEG_ST_N_BAD, (unsigned long long)r[3]);
    }
    atomicAdd(p.stats + B200SEG_ST_N_CORRECT, (unsigned long long)r[2]);
    atomicAdd(p.stats + B200SEG_ST_N_ACC, (unsigned long long)r[4]);
  }
}

// ------------------------------------------------------------------------------------------------
// Backward: grad_z_j = p_j * (g_j - sum_c p_c g_c) [dice, g = dL/dp]  +  k * (p_j - onehot_j) [CE]
template <typename T, int V, int CPT>
__global__ void __launch_bounds__(512) tile_bwd_kernel(const TileParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = p.G, PG = p.PG, C = p.C;
  const int pg = warp / G, g = warp - pg * G;
  const int c0 = g * p.cpg;
  const int c1 = (c0 + p.cpg < C) ? c0 + p.cpg : C;
  const int n = blockIdx.y;
  const long long HW = p.HW;
  constexpr int PXW = 32 * V;

  float* exch_d = reinterpret_cast<float*>(smem_raw);           // [PG][G][PXW] partial dots
  float* exch_k = exch_d + PG * G * PXW;                          // [PG][PXW] CE coefficient
  float* exch_a = exch_k + PG * PXW;                              // [PG][PXW] dice one-hot coefficient
  int* exch_y = reinterpret_cast<int*>(exch_a + PG * PXW);        // [PG][PXW] clamped label

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
    const long long px0 = ((long long)tile * PG + pg) * PXW + (long long)lane * V;
    const bool active = px0 < HW;
    float z[CPT][V];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      if (active && c0 + i < c1) {
        load_vec<T, V>(img + (size_t)(c0 + i) * HW + px0, z[i]);
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) z[i][v] = neg_inf();
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
    // p, and the dense part of sum_c p_c g_c
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
    const int pslot = pg * PXW + lane * V;
    if (g == 0) {
      long long y[V];
      float pwv[V], gpx[V];
#pragma unroll
      for (int v = 0; v < V; ++v) { y[v] = -1; pwv[v] = 1.f; gpx[v] = 1.f; }
      if (active) {
        load_labels<V>(p.labels, p.label_dtype, (size_t)n * HW + px0, y);
        if (want_ce && p.pw) load_vec<float, V>(p.pw + (size_t)n * HW + px0, pwv);
        if (want_ce && p.ce_grad_px) load_vec<float, V>(p.ce_grad_px + (size_t)n * HW + px0, gpx);
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float k = 0.f, da = 0.f;
        int yc = -1;
        if (active) {
          const long long yy = y[v];
          yc = yy < 0 ? 0 : (yy >= (long long)C ? C - 1 : (int)yy);
          if (yy != p.dice_ignore) {
            da = god * __ldg(p.dice_coef + ((size_t)n * C + yc) * 2 + 0);
            const float zy = to_float<T>(img[(size_t)yc * HW + px0 + v]);
            const float py = ex2(fmaf(zy, kLog2e, nl[v]));
            dotp[v] -= da * py;
          }
          const bool valid = (yy != p.ignore_index) && yy >= 0 && yy < (long long)C;
          if (want_ce && valid) k = Gce * pwv[v] * gpx[v] * (p.cw ? __ldg(p.cw + yy) : 1.f);
        }
        exch_k[pslot + v] = k;
        exch_a[pslot + v] = da;
        exch_y[pslot + v] = yc;
      }
    }
    {
      const int slot = (pg * G + g) * PXW + lane * V;
#pragma unroll
      for (int v = 0; v < V; ++v) exch_d[slot + v] = dotp[v];
    }
    group_barrier(pg, 32 * G);
    float kk[V], da[V], sub[V];
    int yc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float t = 0.f;
      for (int gg = 0; gg < G; ++gg) t += exch_d[(pg * G + gg) * PXW + lane * V + v];
      kk[v] = exch_k[pslot + v];
      da[v] = exch_a[pslot + v];
      yc[v] = exch_y[pslot + v];
      sub[v] = kk[v] - t;  // grad = p * (gd + k - dot)
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        if (c0 + i < c1) {
          float gr[V];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float pr = z[i][v];
            float gd;
            if (e2) gd = beta[i] * pr;
            else gd = pr > 0.f ? beta[i] * __powf(pr, p.dice_exponent - 1.f) : 0.f;
            float gv = pr * (gd + sub[v]);
            if (c0 + i == yc[v]) gv -= fmaf(pr, da[v], kk[v]);
            gr[v] = gv;
          }
          store_vec<T, V>(gimg + (size_t)(c0 + i) * HW + px0, gr);
        }
      }
    }
    // the exchange buffers are rewritten next tile: everyone must be done reading
    group_barrier(pg, 32 * G);
  }
}

// ------------------------------------------------------------------------------------------------
static void fill_split(TileParams& p, int V, int cpt) {
  p.G = (p.C + cpt - 1) / cpt;
  p.cpg = (p.C + p.G - 1) / p.G;  // balanced: C=150, cpt=16 -> G=10 x 15 classes
  p.PG = 4 / p.G;
  if (p.PG < 1) p.PG = 1;
  const long long per_tile = (long long)p.PG * 32 * V;
  p.tiles = (int)((p.HW + per_tile - 1) / per_tile);
}

// B200SEG_TILE_CPT={8,16,32} overrides the classes-per-thread heuristic (tuning knob).
static int tile_cpt_override() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200SEG_TILE_CPT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

template <typename T, int V, int CPT> static int launch_tile_fwd(TileParams p, cudaStream_t st) {
  fill_split(p, V, CPT);
  const size_t smem = (size_t)3 * p.PG * p.G * 32 * V * 4 + (size_t)2 * p.PG * p.C * 4;
  int gx = (kSMs * 6 + p.N - 1) / p.N;
  if (gx > p.tiles) gx = p.tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, p.N);
  tile_fwd_kernel<T, V, CPT><<<grid, 32 * p.G * p.PG, smem, st>>>(p);
  count_launch();
  return check_launch("tile_fwd_kernel");
}

template <typename T, int V, int CPT> static int launch_tile_bwd(TileParams p, cudaStream_t st) {
  fill_split(p, V, CPT);
  const size_t smem = (size_t)p.PG * p.G * 32 * V * 4 + (size_t)3 * p.PG * 32 * V * 4;
  int gx = (kSMs * 6 + p.N - 1) / p.N;
  if (gx > p.tiles) gx = p.tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, p.N);
  tile_bwd_kernel<T, V, CPT><<<grid, 32 * p.G * p.PG, smem, st>>>(p);
  count_launch();
  return check_launch("tile_bwd_kernel");
}

#define B200SEG_TILE_LAUNCH(VV, CC) (FWD ? launch_tile_fwd<T, VV, CC>(p, st) : launch_tile_bwd<T, VV, CC>(p, st))
template <typename T, bool FWD> static int tile_pick(const TileParams& p, bool vec, cudaStream_t st) {
  // classes per thread: small C keeps one warp per pixel group; large C splits the class dimension
  // over up to 16 warps (512 threads).
  const int C = p.C;
  int cpt = C <= 8 ? 8 : (C <= 16 ? 16 : (C <= 32 ? 32 : (C <= 256 ? 16 : 32)));
  const int ov = tile_cpt_override();
  if ((ov == 8 || ov == 16 || ov == 32) && (C + ov - 1) / ov <= 16) cpt = ov;
  if (vec) {
    if (cpt == 8) return B200SEG_TILE_LAUNCH(2, 8);
    if (cpt == 16) return B200SEG_TILE_LAUNCH(2, 16);
    return B200SEG_TILE_LAUNCH(2, 32);
  }
  if (cpt == 8) return B200SEG_TILE_LAUNCH(1, 8);
  if (cpt == 16) return B200SEG_TILE_LAUNCH(1, 16);
  return B200SEG_TILE_LAUNCH(1, 32);
}

int tile_fwd_dispatch(const b200seg_loss_desc* d, cudaStream_t st) {
  B200SEG_REQUIRE(d->h == d->H && d->w == d->W, "dice forward needs logits at label resolution (resize first)");
  B200SEG_REQUIRE(d->C <= 512, "dice path supports at most 512 classes (got %d)", d->C);
  B200SEG_REQUIRE(d->dice_part != nullptr, "dice_part workspace is NULL");
  B200SEG_REQUIRE(d->dice_exponent > 0.f, "dice exponent must be > 0");
  TileParams p = {};
  p.logits = d->logits; p.labels = d->labels; p.pw = d->pixel_weight; p.cw = d->ce_class_weight;
  p.lse_out = (d->flags & B200SEG_WANT_LSE) ? d->lse : nullptr;
  p.loss_px = (d->flags & B200SEG_WANT_LOSS_PX) ? d->loss_px : nullptr;
  p.stats = reinterpret_cast<unsigned long long*>(d->stats);
  p.dice_part = d->dice_part;
  p.label_dtype = d->label_dtype;
  p.N = d->N; p.C = d->C; p.HW = (long long)d->H * d->W;
  p.flags = d->flags; p.ignore_index = d->ignore_index;
  p.acc_has_ignore = d->acc_has_ignore; p.acc_ignore = d->acc_ignore_index;
  p.dice_ignore = d->dice_ignore_index; p.dice_exponent = d->dice_exponent; p.lw = d->ce_loss_weight;
  const bool vec = (p.HW % 2 == 0) && aligned16(d->logits) && aligned16(d->labels) &&
                   (!p.pw || aligned16(p.pw)) && (!p.lse_out || aligned16(p.lse_out)) &&
                   (!p.loss_px || aligned16(p.loss_px));
  switch (d->logit_dtype) {
    case B200SEG_F32: return tile_pick<float, true>(p, vec, st);
    case B200SEG_BF16: return tile_pick<__nv_bfloat16, true>(p, vec, st);
    case B200SEG_F16: return tile_pick<__half, true>(p, vec, st);
  }
  set_error("unsupported logit dtype %d", d->logit_dtype);
  return 1;
}

int tile_bwd_dispatch(const b200seg_loss_bwd_desc* d, cudaStream_t st) {
  B200SEG_REQUIRE(d->h == d->H && d->w == d->W, "dice backward needs logits at label resolution");
  B200SEG_REQUIRE(d->C <= 512, "dice path supports at most 512 classes (got %d)", d->C);
  B200SEG_REQUIRE(d->dice_coef != nullptr && d->lse != nullptr, "dice backward needs dice_coef and lse");
  TileParams p = {};
  p.logits = d->logits; p.labels = d->labels; p.pw = d->pixel_weight; p.cw = d->ce_class_weight;
  p.lse_in = d->lse; p.ce_grad_out = d->ce_grad_out; p.ce_grad_px = d->ce_grad_px;
  p.dice_coef = d->dice_coef; p.dice_grad_out = d->dice_grad_out; p.grad = d->grad_logits;
  p.ce_scale_host = d->ce_scale_host; p.ce_use_nvalid = d->ce_use_nvalid;
  p.stats = const_cast<unsigned long long*>(reinterpret_cast<const unsigned long long*>(d->stats));
  p.label_dtype = d->label_dtype;
  p.N = d->N; p.C = d->C; p.HW = (long long)d->H * d->W;
  p.flags = d->flags; p.ignore_index = d->ignore_index;
  p.dice_ignore = d->dice_ignore_index; p.dice_exponent = d->dice_exponent;
  const bool vec = (p.HW % 2 == 0) && aligned16(d->logits) && aligned16(d->labels) && aligned16(d->lse) &&
                   aligned16(d->grad_logits) && (!p.pw || aligned16(p.pw)) && (!p.ce_grad_px || aligned16(p.ce_grad_px));
  switch (d->logit_dtype) {
    case B200SEG_F32: return tile_pick<float, false>(p, vec, st);
    case B200SEG_BF16: return tile_pick<__nv_bfloat16, false>(p, vec, st);
    case B200SEG_F16: return tile_pick<__half, false>(p, vec, st);
  }
  set_error("unsupported logit dtype %d", d->logit_dtype);
  return 1;
}

}  // namespace b200seg
