// Resize-fused soft-max cross-entropy forward+backward for ANY up-sampling ratio and both align_corners settings,
// cell-owner formulation with one THREAD per cell (row group), sm_100a.
//
// Replaces, for logits at lower resolution than the labels (H >= h, W >= w, C <= 32), the chain
// resize (utils/ops.py:7-26) -> cross_entropy (models/losses/cross_entropy_loss.py:23-74) -> accuracy
// (models/losses/accuracy.py:6-61) and the autograd backward of all three, as called from
// models/decode_heads/decode_head.py:261-321, without materialising the (N,C,H,W) tensor in either direction and without
// atomics: ATen's upsample_bilinear2d_backward is an atomicAdd scatter (non-deterministic); here every low-resolution
// logit's gradient is a fixed-order sum.
//
// Geometry (ATen, torch/include/ATen/native/UpSample.h:271-312). Output row Y reads the tap rows (y0, y1) with weight
// ly; y0(Y) is non-decreasing, so the rows split into h + 1 BANDS: band 0 = rows whose source index was clamped to 0
// (align_corners=False only: taps (0,0)), band b = rows with y0 = b - 1, band h = rows with y0 = y1 = h - 1; likewise
// w + 1 RUNS of columns. The pixels of cell (band b, run r) read exactly the 4 low-resolution logits
// (max(b-1,0) | min(b,h-1)) x (max(r-1,0) | min(r,w-1)) per class and scatter their gradient to exactly those 4. Cell
// extents follow from the source-index map itself (any ratio, any align_corners), not from a fixed scale.
//
// Mapping. One thread owns one cell (or 1/RG of its rows). For a row of the cell and a class c the interpolated logits
// of the run's pixels are z_j = L_c + lx_j D_c with lx_j EQUALLY SPACED (step = the horizontal scale), so their
// exponentials are a geometric progression e_j = E_0 R^j: two MUFU.EX2 per class and row, then one FMUL per class-pixel
// (product tree of depth 3) instead of FFMA + MUFU. Forward sweep over the classes: chain, sum, max. Per-pixel scalars
// (label, its logit, loss, accuracy, 1 / sum). Backward sweep: the chain again, two FFMA per class-pixel into the row's
// horizontal corner sums, then one read-modify-write of the class's 4 corner sums in the thread's private shared-memory
// column. No cross-lane exchange per pixel at all. The cell's corner logits live in shared memory, pre-scaled by log2 e
// and offset by the cell's maximum (every interpolated logit is a convex combination of its corners, so z <= 0).
// Each cell's sums go once to PB[n][c][band][run] (float4); up_combine_kernel (loss_up.cu) adds the 4 cells around every
// logit. A chunk falls back to DIRECT evaluation (FFMA + MUFU per class-pixel, exact per-pixel maxima) when the chain is
// not applicable: a partially filled chunk at a ratio below PXC - 1, a class more than 100 log2-units below the cell maximum at either tap
// (its chain would leave the normal range), or an underflowing sum.
//
// Top-1: the label's class is the arg-max iff its exponential reaches the pixel's maximum exponential up to the chain's
// rounding (2^-19 relative): exact and near ties count as correct (torch.topk's choice among ties is unspecified).
// Bound: instruction issue (~9 issue slots per class-pixel, 0.5 MUFU); HBM traffic is the label map.
// Algorithmic bytes per launch: 2*N*C*h*w*s + N*H*W*L.
#pragma once
#include "common.cuh"
#include "loss_upcell.cuh"   // RawLabel / decode_label / lg2 / pixel_weight / kLn2

#ifndef B200SEG_UPGEN_UNROLL
#define B200SEG_UPGEN_UNROLL 4   // classes in flight per thread in the class sweeps (2, 3, 4 measure 81.4 / 80.7 / 80.4 us at config 2)
#endif

namespace b200seg {

constexpr int kUpgenUnroll = B200SEG_UPGEN_UNROLL;

struct UpGenParams {
  const void* logits;
  const void* labels;
  const float* pw;
  const float* cw;
  unsigned long long* stats;
  float* pb;
  float* lse2;        // (N,H,W) per-pixel log2-sum-exp of the interpolated logits: written by the forward-only launch of
                      // the class-tiled plan (C > 32), read by its backward launches; NULL otherwise
  int label_dtype, label_bytes;
  int has_w;
  int N, C, h, w, H, W;
  int ac;
  float sh, sw;       // ATen's area_pixel_compute_scale for rows / columns
  int RG, logRG;      // threads per cell (row groups)
  long long cells;    // N * (h + 1) * (w + 1)
  int ignore32, acc_has_ignore, acc_ignore32;
};

// ATen's source index of output position `dst` (fp32, area_pixel_compute_source_index), before clamping to >= 0
__device__ __forceinline__ float up_raw_src(float scale, int dst, bool ac) {
  return ac ? scale * (float)dst : scale * ((float)dst + 0.5f) - 0.5f;
}
// band key of an output position: 0 = clamped below 0, k + 1 = source floor k (clamped to in - 1)
__device__ __forceinline__ int up_key(float scale, int dst, int in, bool ac) {
  const float s = up_raw_src(scale, dst, ac);
  if (s < 0.f) return 0;
  const int i = (int)s;
  return (i < in - 1 ? i : in - 1) + 1;
}
// first output position in [0, out] whose key is >= b (keys are non-decreasing in dst)
__device__ __forceinline__ int up_band_start(float scale, int b, int in, int out, bool ac) {
  if (b <= 0) return 0;
  if (b > in || !(scale > 0.f)) return out;
  const float est = ac ? ((float)(b - 1) / scale) : (((float)(b - 1) + 0.5f) / scale - 0.5f);
  int d = (int)fminf(fmaxf(ceilf(est), 0.f), (float)out);
  while (d > 0 && up_key(scale, d - 1, in, ac) >= b) --d;
  while (d < out && up_key(scale, d, in, ac) < b) ++d;
  return d;
}
// lambda of a position inside a regular band (source floor k = b - 1)
__device__ __forceinline__ float up_lambda(float scale, int dst, int k, bool ac) {
  float l = up_raw_src(scale, dst, ac) - (float)k;
  return l < 0.f ? 0.f : (l > 1.f ? 1.f : l);
}

// e[j] = E0 * R^j as a product tree of depth <= 3; every use of an element (class sweep, label) runs the SAME products
template <int PXC> __device__ __forceinline__ void up_chain(float E0, float R, float (&e)[PXC]) {
  static_assert(PXC == 4 || PXC == 8, "chunk width");
  const float R2 = R * R;
  e[0] = E0;
  e[1] = E0 * R;
  e[2] = E0 * R2;
  e[3] = e[1] * R2;
  if constexpr (PXC == 8) {
    const float R4 = R2 * R2;
    e[4] = E0 * R4;
    e[5] = e[1] * R4;
    e[6] = e[2] * R4;
    e[7] = e[3] * R4;
  }
}

template <int THR> constexpr size_t upgen_smem_bytes(int C, int logRG, bool grad) {
  return (size_t)C * (THR >> logRG) * 16 + (grad ? (size_t)C * THR * 16 : 0);
}

template <typename T, int PXC, bool GRAD, int LK, int THR>
__global__ void __launch_bounds__(THR) up_gen_kernel(const UpGenParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int C = p.C;
  const int RG = p.RG;
  const int cpc = THR >> p.logRG;                                  // cells per CTA
  float4* CORN = reinterpret_cast<float4*>(smem_raw);             // [C][cpc]  (a, b, da, db) per class of the cell
  float4* OH = CORN + (size_t)C * cpc;                            // [C][THR]  private corner sums (GRAD)
  const int rg = tid & (RG - 1);
  const int cell = tid >> p.logRG;
  const unsigned cid_raw = blockIdx.x * (unsigned)cpc + (unsigned)cell;
  const bool cell_ok = cid_raw < (unsigned)p.cells;
  const unsigned cid = cell_ok ? cid_raw : (unsigned)p.cells - 1u;
  const unsigned t0 = cid / (unsigned)(p.w + 1);
  const int r = (int)(cid - t0 * (unsigned)(p.w + 1));
  const int n = (int)(t0 / (unsigned)(p.h + 1));
  const int b = (int)(t0 - (unsigned)n * (unsigned)(p.h + 1));
  const bool ac = p.ac != 0;

  // ---- the cell's 4 corner logits: the RG threads of the cell split the classes; scaled by log2 e and offset by the
  // cell's maximum once the latter is known
  float M2cell = 0.f;
  {
    const int plane = p.h * p.w;
    const T* pl = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * (size_t)plane;
    const int ya = b - 1 < 0 ? 0 : b - 1, yb = b > p.h - 1 ? p.h - 1 : b;
    const int xa = r - 1 < 0 ? 0 : r - 1, xb = r > p.w - 1 ? p.w - 1 : r;
    const int o00 = ya * p.w + xa, o01 = ya * p.w + xb, o10 = yb * p.w + xa, o11 = yb * p.w + xb;
    float M = -3.0e38f;
    for (int c = rg; c < C; c += RG) {
      const T* q = pl + (size_t)c * plane;
      const float v00 = to_float<T>(__ldg(q + o00)), v01 = to_float<T>(__ldg(q + o01));
      const float v10 = to_float<T>(__ldg(q + o10)), v11 = to_float<T>(__ldg(q + o11));
      M = fmaxf(fmaxf(M, fmaxf(v00, v01)), fmaxf(v10, v11));
      CORN[c * cpc + cell] = make_float4(v00, v01, v10, v11);
    }
    for (int off = 1; off < RG; off <<= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, off));
    const float nM2 = -M * kLog2e;
    M2cell = M * kLog2e;
    for (int c = rg; c < C; c += RG) {
      const float4 v = CORN[c * cpc + cell];
      CORN[c * cpc + cell] = make_float4(fmaf(v.x, kLog2e, nM2), fmaf(v.y, kLog2e, nM2), (v.z - v.x) * kLog2e, (v.w - v.y) * kLog2e);
    }
    if constexpr (GRAD) {
      for (int c = 0; c < C; ++c) OH[c * THR + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
  }
  const float4* corn = CORN + cell;

  // ---- extents of the cell and this thread's rows
  const int Yb0 = up_band_start(p.sh, b, p.h, p.H, ac), Yb1 = up_band_start(p.sh, b + 1, p.h, p.H, ac);
  const int X0 = up_band_start(p.sw, r, p.w, p.W, ac), X1 = up_band_start(p.sw, r + 1, p.w, p.W, ac);
  const int rows_per = (Yb1 - Yb0 + RG - 1) >> p.logRG;
  const int Yr0 = Yb0 + rg * rows_per;
  const int Yr1 = cell_ok ? min(Yb1, Yr0 + rows_per) : Yr0;
  // regular run: lambda advances by the scale. Clamped runs / bands read one tap only: the weight is put on the tap the
  // combine step reads for that cell — the HIGH tap of band 0 / run 0 (row 0 / column 0), the LOW tap of band h / run w.
  const bool xreg = (r > 0 && r < p.w);
  const float sx = xreg ? p.sw : 0.f;
  const float lx0 = xreg ? (X0 < X1 ? up_lambda(p.sw, X0, r - 1, ac) : 0.f) : (r == 0 ? 1.f : 0.f);
  const bool yreg = (b > 0 && b < p.h);
  const float ly_clamped = (b == 0) ? 1.f : 0.f;

  float loss_acc = 0.f;
  int n_valid = 0, n_correct = 0, n_bad = 0, n_acc = 0;
  const int dt = p.label_dtype;
  const int lb = LK == 0 ? 8 : (LK == 1 ? 1 : p.label_bytes);
  const size_t img_px = (size_t)n * p.H * p.W;
  const char* labimg = reinterpret_cast<const char*>(p.labels) + img_px * lb;

#pragma unroll 1
  for (int Y = Yr0; Y < Yr1; ++Y) {
    const float ly = yreg ? up_lambda(p.sh, Y, b - 1, ac) : ly_clamped;
    const unsigned roff = (unsigned)Y * (unsigned)p.W;
#pragma unroll 1
    for (int Xc = X0; Xc < X1; Xc += PXC) {
      const int npx = min(PXC, X1 - Xc);
      // ---- labels of the chunk, issued before the class sweep (consumed after it)
      RawLabel raw[PXC];
#pragma unroll
      for (int j = 0; j < PXC; ++j) raw[j] = load_raw_label<LK>(labimg, dt, roff + (unsigned)(Xc + min(j, npx - 1)));
      const float lam0 = fmaf((float)(Xc - X0), sx, lx0);
      float s[PXC], m[PXC], moff[PXC];
#pragma unroll
      for (int j = 0; j < PXC; ++j) { s[j] = 0.f; m[j] = 0.f; moff[j] = 0.f; }
      // a partially filled chunk still takes the chain when the lambdas of its masked tail stay <= 2 (then, with every
      // tap >= -100 below the maximum, the tail's exponentials stay finite; their weights a[j] are 0)
      bool fast = (npx == PXC) || (sx * (float)(PXC - 1) <= 1.f);
      if (fast) {
        // ---- forward sweep, geometric chain
        float minend = 0.f;
#pragma unroll kUpgenUnroll
        for (int c = 0; c < C; ++c) {
          const float4 q = corn[c * cpc];
          const float L2 = fmaf(ly, q.z, q.x), R2 = fmaf(ly, q.w, q.y);
          const float D2 = R2 - L2;
          minend = fminf(minend, fminf(L2, R2));
          float e[PXC];
          up_chain<PXC>(ex2(fmaf(lam0, D2, L2)), ex2(D2 * sx), e);
#pragma unroll
          for (int j = 0; j < PXC; ++j) { s[j] += e[j]; m[j] = fmaxf(m[j], e[j]); }
        }
        // taps within 100 log2-units of the cell maximum: the chain stays in the normal range and the Horner partial sums of
        // the backward sweep (<= weight / E0) stay finite
        bool okc = minend >= -100.f;
#pragma unroll
        for (int j = 0; j < PXC; ++j) okc = okc && (j >= npx || ((s[j] > 1e-30f) && (s[j] < 3.0e38f)));
        fast = okc;
      }
      if (!fast) {
        // ---- direct evaluation against exact per-pixel maxima (lambda of the masked tail pixels = the last valid one)
#pragma unroll
        for (int j = 0; j < PXC; ++j) { s[j] = 0.f; m[j] = 0.f; moff[j] = -3.0e38f; }
        for (int c = 0; c < C; ++c) {
          const float4 q = corn[c * cpc];
          const float L2 = fmaf(ly, q.z, q.x), D2 = fmaf(ly, q.w, q.y) - L2;
#pragma unroll
          for (int j = 0; j < PXC; ++j) moff[j] = fmaxf(moff[j], fmaf(fmaf((float)min(j, npx - 1), sx, lam0), D2, L2));
        }
        for (int c = 0; c < C; ++c) {
          const float4 q = corn[c * cpc];
          const float L2 = fmaf(ly, q.z, q.x), D2 = fmaf(ly, q.w, q.y) - L2;
#pragma unroll
          for (int j = 0; j < PXC; ++j) {
            const float ev = ex2(fmaf(fmaf((float)min(j, npx - 1), sx, lam0), D2, L2) - moff[j]);
            s[j] += ev;
            m[j] = fmaxf(m[j], ev);
          }
        }
      }

      // ---- per-pixel scalars
      float a[PXC], bl[PXC];
#pragma unroll
      for (int j = 0; j < PXC; ++j) {
        const bool ok = cell_ok && j < npx;
        const int ydec = decode_label<LK>(raw[j], dt);
        const int yy = ok ? ydec : p.ignore32;
        const bool ign = (yy == p.ignore32);
        const bool inr = (unsigned)yy < (unsigned)C;
        const int yc = inr && !ign ? yy : 0;
        const bool use = ok && inr && !ign;
        const bool acc_ok = ok && (p.acc_has_ignore ? (yy != p.acc_ignore32) : true);
        float wt = use ? 1.f : 0.f;
        if (p.has_w) wt = pixel_weight(p.cw, p.pw, use, yc, img_px + (size_t)(roff + (unsigned)(Xc + min(j, npx - 1))));
        n_valid += (ok && !ign);
        n_bad += (ok && !ign && !inr);
        n_acc += acc_ok;
        const float lamj = fmaf((float)(fast ? j : min(j, npx - 1)), sx, lam0);
        // the label's interpolated logit and exponential: the operations of the class sweep on the label's corners
        const float4 q = corn[yc * cpc];
        const float L2 = fmaf(ly, q.z, q.x), R2 = fmaf(ly, q.w, q.y);
        const float D2 = R2 - L2;
        const float zy2 = fmaf(lamj, D2, L2);
        // the label is the arg-max iff its exponential reaches the pixel's maximum exponential; the maximum comes out of
        // the product chain (relative error <= 2^-20: four ex2.approx factors), so the test allows 2^-19 — an exact or
        // near tie with another class counts as correct for the label (torch.topk's choice among ties is unspecified)
        const float ey = ex2(zy2 - moff[j]);
        const float lse2_rel = moff[j] + lg2(s[j]);
        loss_acc = fmaf(wt, lse2_rel - zy2, loss_acc);
        n_correct += (acc_ok && inr && !ign && ey >= m[j] * 0.99999809265f);
        if constexpr (!GRAD) {
          if (p.lse2 && ok) p.lse2[img_px + (size_t)(roff + (unsigned)(Xc + j))] = M2cell + lse2_rel;
        }
        if constexpr (GRAD) {
          a[j] = wt * fast_rcp(s[j]);
          bl[j] = a[j] * lamj;
          // one-hot term of the pixel into the label class's corner sums
          const float u = wt * lamj, v = wt - u;
          float4* oh = OH + yc * THR + tid;
          float4 o = *oh;
          o.x = fmaf(ly - 1.f, v, o.x);
          o.y = fmaf(ly - 1.f, u, o.y);
          o.z = fmaf(-ly, v, o.z);
          o.w = fmaf(-ly, u, o.w);
          *oh = o;
        }
      }

      // ---- backward sweep: horizontal corner sums of wt * softmax per class, folded into the class's corner sums
      if constexpr (GRAD) {
        const float ly0 = 1.f - ly;
        if (fast) {
#pragma unroll kUpgenUnroll
          for (int c = 0; c < C; ++c) {
            const float4 q = corn[c * cpc];
            const float L2 = fmaf(ly, q.z, q.x), D2 = fmaf(ly, q.w, q.y) - L2;
            // sum_j a_j E0 R^j = E0 * A(R): the row's weighted sums are two polynomials in R (coefficients a_j >= 0 and
            // a_j lambda_j >= 0, no cancellation) evaluated by Horner's rule — 2 (PXC - 1) FFMA instead of the product chain
            // plus 2 PXC FFMA
            const float E0 = ex2(fmaf(lam0, D2, L2)), R = ex2(D2 * sx);
            float pa = a[PXC - 1], pb = bl[PXC - 1];
#pragma unroll
            for (int j = PXC - 2; j >= 0; --j) { pa = fmaf(pa, R, a[j]); pb = fmaf(pb, R, bl[j]); }
            const float gs = E0 * pa, gb = E0 * pb;
            const float ga = gs - gb;
            float4* oh = OH + c * THR + tid;
            float4 o = *oh;
            o.x = fmaf(ly0, ga, o.x);
            o.y = fmaf(ly0, gb, o.y);
            o.z = fmaf(ly, ga, o.z);
            o.w = fmaf(ly, gb, o.w);
            *oh = o;
          }
        } else {
          for (int c = 0; c < C; ++c) {
            const float4 q = corn[c * cpc];
            const float L2 = fmaf(ly, q.z, q.x), D2 = fmaf(ly, q.w, q.y) - L2;
            float gs = 0.f, gb = 0.f;
#pragma unroll
            for (int j = 0; j < PXC; ++j) {
              const float ev = ex2(fmaf(fmaf((float)min(j, npx - 1), sx, lam0), D2, L2) - moff[j]);
              gs = fmaf(ev, a[j], gs);
              gb = fmaf(ev, bl[j], gb);
            }
            const float ga = gs - gb;
            float4* oh = OH + c * THR + tid;
            float4 o = *oh;
            o.x = fmaf(ly0, ga, o.x);
            o.y = fmaf(ly0, gb, o.y);
            o.z = fmaf(ly, ga, o.z);
            o.w = fmaf(ly, gb, o.w);
            *oh = o;
          }
        }
      }
    }
  }

  if constexpr (GRAD) {
    // the RG private columns of a cell are added in a fixed order: thread rg takes the classes c = rg (mod RG) and reads
    // the RG columns straight from shared memory (no shuffle tree), then writes those classes of the cell
    __syncwarp();
    if (cell_ok) {
      const int col0 = tid - rg;
      for (int c = rg; c < C; c += RG) {
        float4 o = OH[c * THR + col0];
        for (int k = 1; k < RG; ++k) {
          const float4 t = OH[c * THR + col0 + k];
          o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
        }
        float4* dst = reinterpret_cast<float4*>(p.pb) + (((size_t)n * C + c) * (p.h + 1) + b) * (p.w + 1) + r;
        *dst = o;
      }
    }
  }
  cta_flush_stats(loss_acc * kLn2, n_valid, n_correct, n_bad, n_acc, p.stats);
}

// ------------------------------------------------------------------------------------------------ class-tiled backward
// C > 32: a thread's private corner sums for ALL classes do not fit shared memory (16 C bytes per thread), so the plan is
// split: one forward-only launch of up_gen_kernel over all classes (loss, accuracy, and the per-pixel log2-sum-exp into
// `lse2`), then this kernel once per tile of kUpTile classes (grid.y): soft-max probabilities of the tile's classes from
// the saved lse2 — p = 2^(z2 - lse2) — the same chain / corner-sum machinery, the tile's slice of PB. Deterministic.
constexpr int kUpTile = 32;

template <typename T, int PXC, int LK, int THR>
__global__ void __launch_bounds__(THR) up_gen_bwd_tile_kernel(const UpGenParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int C = p.C;
  const int RG = p.RG;
  const int cpc = THR >> p.logRG;
  float4* CORN = reinterpret_cast<float4*>(smem_raw);             // [kUpTile][cpc]
  float4* OH = CORN + (size_t)kUpTile * cpc;                      // [kUpTile][THR]
  const int c0 = blockIdx.y * kUpTile;
  const int ct = min(kUpTile, C - c0);
  const int rg = tid & (RG - 1);
  const int cell = tid >> p.logRG;
  const unsigned cid_raw = blockIdx.x * (unsigned)cpc + (unsigned)cell;
  const bool cell_ok = cid_raw < (unsigned)p.cells;
  const unsigned cid = cell_ok ? cid_raw : (unsigned)p.cells - 1u;
  const unsigned t0 = cid / (unsigned)(p.w + 1);
  const int r = (int)(cid - t0 * (unsigned)(p.w + 1));
  const int n = (int)(t0 / (unsigned)(p.h + 1));
  const int b = (int)(t0 - (unsigned)n * (unsigned)(p.h + 1));
  const bool ac = p.ac != 0;

  float M2t;
  bool chain_ok;
  {
    const int plane = p.h * p.w;
    const T* pl = reinterpret_cast<const T*>(p.logits) + ((size_t)n * C + c0) * (size_t)plane;
    const int ya = b - 1 < 0 ? 0 : b - 1, yb = b > p.h - 1 ? p.h - 1 : b;
    const int xa = r - 1 < 0 ? 0 : r - 1, xb = r > p.w - 1 ? p.w - 1 : r;
    const int o00 = ya * p.w + xa, o01 = ya * p.w + xb, o10 = yb * p.w + xa, o11 = yb * p.w + xb;
    float M = -3.0e38f, mn = 3.0e38f;
    for (int c = rg; c < ct; c += RG) {
      const T* q = pl + (size_t)c * plane;
      const float v00 = to_float<T>(__ldg(q + o00)), v01 = to_float<T>(__ldg(q + o01));
      const float v10 = to_float<T>(__ldg(q + o10)), v11 = to_float<T>(__ldg(q + o11));
      M = fmaxf(fmaxf(M, fmaxf(v00, v01)), fmaxf(v10, v11));
      mn = fminf(fminf(mn, fminf(v00, v01)), fminf(v10, v11));
      CORN[c * cpc + cell] = make_float4(v00, v01, v10, v11);
    }
    for (int off = 1; off < RG; off <<= 1) {
      M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, off));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    }
    M2t = M * kLog2e;
    // the chain (and the per-pixel factor 2^(M2t - lse2) <= 2^(M2t - min)) stays in the normal range
    chain_ok = (M - mn) * kLog2e <= 100.f;
    const float nM2 = -M2t;
    for (int c = rg; c < ct; c += RG) {
      const float4 v = CORN[c * cpc + cell];
      CORN[c * cpc + cell] = make_float4(fmaf(v.x, kLog2e, nM2), fmaf(v.y, kLog2e, nM2), (v.z - v.x) * kLog2e, (v.w - v.y) * kLog2e);
    }
    for (int c = 0; c < ct; ++c) OH[c * THR + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
  }
  const float4* corn = CORN + cell;

  const int Yb0 = up_band_start(p.sh, b, p.h, p.H, ac), Yb1 = up_band_start(p.sh, b + 1, p.h, p.H, ac);
  const int X0 = up_band_start(p.sw, r, p.w, p.W, ac), X1 = up_band_start(p.sw, r + 1, p.w, p.W, ac);
  const int rows_per = (Yb1 - Yb0 + RG - 1) >> p.logRG;
  const int Yr0 = Yb0 + rg * rows_per;
  const int Yr1 = cell_ok ? min(Yb1, Yr0 + rows_per) : Yr0;
  const bool xreg = (r > 0 && r < p.w);
  const float sx = xreg ? p.sw : 0.f;
  const float lx0 = xreg ? (X0 < X1 ? up_lambda(p.sw, X0, r - 1, ac) : 0.f) : (r == 0 ? 1.f : 0.f);
  const bool yreg = (b > 0 && b < p.h);
  const float ly_clamped = (b == 0) ? 1.f : 0.f;
  const int dt = p.label_dtype;
  const int lb = LK == 0 ? 8 : (LK == 1 ? 1 : p.label_bytes);
  const size_t img_px = (size_t)n * p.H * p.W;
  const char* labimg = reinterpret_cast<const char*>(p.labels) + img_px * lb;
  const float* lseimg = p.lse2 + img_px;

#pragma unroll 1
  for (int Y = Yr0; Y < Yr1; ++Y) {
    const float ly = yreg ? up_lambda(p.sh, Y, b - 1, ac) : ly_clamped;
    const float ly0 = 1.f - ly;
    const unsigned roff = (unsigned)Y * (unsigned)p.W;
#pragma unroll 1
    for (int Xc = X0; Xc < X1; Xc += PXC) {
      const int npx = min(PXC, X1 - Xc);
      RawLabel raw[PXC];
      float off2[PXC];
#pragma unroll
      for (int j = 0; j < PXC; ++j) {
        const unsigned px = roff + (unsigned)(Xc + min(j, npx - 1));
        raw[j] = load_raw_label<LK>(labimg, dt, px);
        off2[j] = M2t - __ldg(lseimg + px);           // z2_rel + off2 = z2 - lse2 <= 0
      }
      const float lam0 = fmaf((float)(Xc - X0), sx, lx0);
      const bool fast = chain_ok && ((npx == PXC) || (sx * (float)(PXC - 1) <= 1.f));
      float a[PXC], bl[PXC], wts[PXC];
#pragma unroll
      for (int j = 0; j < PXC; ++j) {
        const bool ok = cell_ok && j < npx;
        const int ydec = decode_label<LK>(raw[j], dt);
        const int yy = ok ? ydec : p.ignore32;
        const bool ign = (yy == p.ignore32);
        const bool inr = (unsigned)yy < (unsigned)C;
        const int yc = inr && !ign ? yy : 0;
        const bool use = ok && inr && !ign;
        float wt = use ? 1.f : 0.f;
        if (p.has_w) wt = pixel_weight(p.cw, p.pw, use, yc, img_px + (size_t)(roff + (unsigned)(Xc + min(j, npx - 1))));
        const float lamj = fmaf((float)(fast ? j : min(j, npx - 1)), sx, lam0);
        wts[j] = wt;
        a[j] = fast ? wt * ex2(off2[j]) : wt;          // direct evaluation folds off2 into the exponent instead
        bl[j] = a[j] * lamj;
        const int yl = yc - c0;
        if (use && (unsigned)yl < (unsigned)ct) {     // one-hot term, for labels of this tile
          const float u = wt * lamj, v = wt - u;
          float4* oh = OH + yl * THR + tid;
          float4 o = *oh;
          o.x = fmaf(ly - 1.f, v, o.x);
          o.y = fmaf(ly - 1.f, u, o.y);
          o.z = fmaf(-ly, v, o.z);
          o.w = fmaf(-ly, u, o.w);
          *oh = o;
        }
      }
      if (fast) {
#pragma unroll kUpgenUnroll
        for (int c = 0; c < ct; ++c) {
          const float4 q = corn[c * cpc];
          const float L2 = fmaf(ly, q.z, q.x), D2 = fmaf(ly, q.w, q.y) - L2;
          const float E0 = ex2(fmaf(lam0, D2, L2)), R = ex2(D2 * sx);     // Horner form of sum_j a_j E0 R^j (see up_gen_kernel)
          float pa = a[PXC - 1], pb = bl[PXC - 1];
#pragma unroll
          for (int j = PXC - 2; j >= 0; --j) { pa = fmaf(pa, R, a[j]); pb = fmaf(pb, R, bl[j]); }
          const float gs = E0 * pa, gb = E0 * pb;
          const float ga = gs - gb;
          float4* oh = OH + c * THR + tid;
          float4 o = *oh;
          o.x = fmaf(ly0, ga, o.x);
          o.y = fmaf(ly0, gb, o.y);
          o.z = fmaf(ly, ga, o.z);
          o.w = fmaf(ly, gb, o.w);
          *oh = o;
        }
      } else {
        for (int c = 0; c < ct; ++c) {
          const float4 q = corn[c * cpc];
          const float L2 = fmaf(ly, q.z, q.x), D2 = fmaf(ly, q.w, q.y) - L2;
          float gs = 0.f, gb = 0.f;
#pragma unroll
          for (int j = 0; j < PXC; ++j) {
            const float ev = wts[j] != 0.f ? ex2(fmaf(fmaf((float)min(j, npx - 1), sx, lam0), D2, L2) + off2[j]) : 0.f;
            gs = fmaf(ev, a[j], gs);
            gb = fmaf(ev, bl[j], gb);
          }
          const float ga = gs - gb;
          float4* oh = OH + c * THR + tid;
          float4 o = *oh;
          o.x = fmaf(ly0, ga, o.x);
          o.y = fmaf(ly0, gb, o.y);
          o.z = fmaf(ly, ga, o.z);
          o.w = fmaf(ly, gb, o.w);
          *oh = o;
        }
      }
    }
  }
  __syncwarp();
  if (cell_ok) {
    const int col0 = tid - rg;
    for (int c = rg; c < ct; c += RG) {
      float4 o = OH[c * THR + col0];
      for (int k = 1; k < RG; ++k) {
        const float4 t = OH[c * THR + col0 + k];
        o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
      }
      float4* dst = reinterpret_cast<float4*>(p.pb) + (((size_t)n * C + c0 + c) * (p.h + 1) + b) * (p.w + 1) + r;
      *dst = o;
    }
  }
}

// resident CTAs per SM from the shared-memory footprint (228 KB per SM, 1 KB reserved per CTA) and the register file
template <int THR> static int upgen_resident_warps(int C, int logRG, bool grad) {
  const size_t smem = upgen_smem_bytes<THR>(C, logRG, grad) + 1024 + 640;
  long long ctas = (long long)(228 * 1024) / (long long)smem;
  const long long by_regs = 65536 / (128 * THR);   // assume <= 128 registers per thread
  if (ctas > by_regs) ctas = by_regs;
  if (ctas > 32) ctas = 32;
  if (ctas < 1) ctas = 1;
  return (int)ctas * (THR / 32);
}

template <typename T, int PXC, bool GRAD, int LK> static int launch_upgen_lk(UpGenParams p, cudaStream_t st) {
  constexpr int THR = 32;
  if (p.C > kUpTile) {
    // class-tiled plan: the corner logits of ALL classes sit in shared memory per CELL (16 C bytes), so 4 threads share a cell
    p.logRG = 2;
    p.RG = 4;
    const long long cells_per_cta = THR >> p.logRG;
    const long long grid = (p.cells + cells_per_cta - 1) / cells_per_cta;
    {
      const size_t smem = upgen_smem_bytes<THR>(p.C, p.logRG, false);
      auto k = up_gen_kernel<T, PXC, false, LK, THR>;
      if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(k), (int)smem)) return e;
      k<<<(unsigned)grid, THR, smem, st>>>(p);
      count_launch();
      if (int e = check_launch("up_gen_kernel")) return e;
    }
    if (GRAD) {
      const size_t smem = upgen_smem_bytes<THR>(kUpTile, p.logRG, true);
      auto k = up_gen_bwd_tile_kernel<T, PXC, LK, THR>;
      if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(k), (int)smem)) return e;
      dim3 g2((unsigned)grid, (unsigned)((p.C + kUpTile - 1) / kUpTile));
      k<<<g2, THR, smem, st>>>(p);
      count_launch();
      return check_launch("up_gen_bwd_tile_kernel");
    }
    return 0;
  }
  p.lse2 = nullptr;
  // row groups: enough threads for >= 2 waves of resident warps, and the fewest idle slots in the last wave
  int best = 0;
  double best_eff = -1.0;
  const int max_rows = (p.H + p.h - 1) / p.h + 1;
  for (int lg = 0; lg <= 2; ++lg) {
    if (lg > 0 && (max_rows >> lg) < 2) break;
    const int warps_res = upgen_resident_warps<THR>(p.C, lg, GRAD) * kSMs;
    const double warps = (double)p.cells * (1 << lg) / 32.0;
    const double waves = warps / warps_res;
    const double eff = waves / (double)(long long)(waves + 0.999999);      // filled fraction of the waves
    const double score = eff - 0.02 * lg - (waves < 1.0 ? 1.0 - waves : 0.0);
    if (score > best_eff) { best_eff = score; best = lg; }
  }
  p.logRG = best;
  p.RG = 1 << best;
  const size_t smem = upgen_smem_bytes<THR>(p.C, p.logRG, GRAD);
  auto k = up_gen_kernel<T, PXC, GRAD, LK, THR>;
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(k), (int)smem)) return e;
  const long long cells_per_cta = THR >> p.logRG;
  const long long grid = (p.cells + cells_per_cta - 1) / cells_per_cta;
  k<<<(unsigned)grid, THR, smem, st>>>(p);
  count_launch();
  return check_launch("up_gen_kernel");
}

template <typename T, int PXC, bool GRAD> static int launch_upgen_px(const UpGenParams& p, cudaStream_t st) {
  if (p.label_dtype == B200SEG_L_I64) return launch_upgen_lk<T, PXC, GRAD, 0>(p, st);
  if (p.label_dtype == B200SEG_L_U8) return launch_upgen_lk<T, PXC, GRAD, 1>(p, st);
  return launch_upgen_lk<T, PXC, GRAD, 2>(p, st);
}

template <typename T> int upgen_run(const b200seg_loss_desc* f, float* pb, bool grad, cudaStream_t st) {
  UpGenParams p;
  p.logits = f->logits; p.labels = f->labels; p.pw = f->pixel_weight; p.cw = f->ce_class_weight;
  p.stats = reinterpret_cast<unsigned long long*>(f->stats);
  p.pb = pb;
  // workspace layout: PB (N,C,h+1,w+1) float4, then — class-tiled plan only — the (N,H,W) per-pixel log2-sum-exp
  p.lse2 = (grad && f->C > kUpTile) ? pb + (size_t)f->N * f->C * (f->h + 1) * (f->w + 1) * 4 : nullptr;
  p.label_dtype = f->label_dtype; p.label_bytes = label_bytes(f->label_dtype);
  p.has_w = (p.cw != nullptr) || (p.pw != nullptr);
  p.N = f->N; p.C = f->C; p.h = f->h; p.w = f->w; p.H = f->H; p.W = f->W;
  p.ac = f->align_corners != 0;
  p.sh = resize_scale(f->h, f->H, p.ac != 0);
  p.sw = resize_scale(f->w, f->W, p.ac != 0);
  p.RG = 1; p.logRG = 0;
  p.cells = (long long)f->N * (f->h + 1) * (f->w + 1);
  auto fit32 = [](long long v) { return (v >= -2147483647LL && v <= 2147483647LL) ? (int)v : kNeverLabel; };
  p.ignore32 = fit32(f->ignore_index); p.acc_has_ignore = f->acc_has_ignore; p.acc_ignore32 = fit32(f->acc_ignore_index);
  if (p.cells == 0) return 0;
  B200SEG_REQUIRE(p.cells < (1LL << 31) && (long long)f->H * f->W < (1LL << 31), "loss_fused: problem too large for 32-bit cell / pixel indices");
  // chunk width: 8 pixels when a run is at least ~6 pixels wide, else 4
  const bool wide = (long long)f->W >= 6LL * f->w;
  if (grad) return wide ? launch_upgen_px<T, 8, true>(p, st) : launch_upgen_px<T, 4, true>(p, st);
  return wide ? launch_upgen_px<T, 8, false>(p, st) : launch_upgen_px<T, 4, false>(p, st);
}

}  // namespace b200seg
