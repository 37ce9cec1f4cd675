// Resize-fused soft-max cross-entropy forward+backward for ANY up-sampling ratio and both align_corners settings,
// cell-owner formulation with one THREAD per cell (row group), sm_100a.
//
// Replaces, for logits at lower resolution than the labels (H >= h, W >= w), the chain
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
// exponentials are a geometric progression e_j = E_0 R^j: two MUFU.EX2 per class and row, then products instead of
// FFMA + MUFU per class-pixel. Issue slots are one of the two units the kernel sits on, so the class sweeps run on
// Blackwell's packed fp32 pipe forms (FMUL2 / FADD2 / FFMA2: two pixels per issue slot) and the 3-input FMNMX3:
//   forward sweep   per class: chain of PXC exponentials (3 FMUL + 3 FMUL2), PXC/2 FADD2 into the pixel sums, and — two
//                   classes at a time — PXC FMNMX3 into the pixel maxima                      (~23 issue slots / class)
//   per-pixel pass  label, its logit, loss, top-1, 1 / sum; one-hot terms folded once per RUN of equal labels
//                                                                                             (~47 issue slots / pixel)
//   backward sweep  per class: sum_j a_j E_0 R^j and sum_j a_j lambda_j E_0 R^j as ONE packed Horner recurrence
//                   (PXC - 1 FFMA2), then one read-modify-write of the class's 4 corner sums in the thread's private
//                   shared-memory column (2 FFMA2)                                            (~21 issue slots / class)
// No cross-lane exchange per pixel at all. The cell's corner logits live in shared memory, pre-scaled by log2 e and
// offset by the cell's maximum (every interpolated logit is a convex combination of its corners, so z <= 0).
// Each cell's sums go once to PB[n][c][band][run] (float4); up_combine_kernel (loss_up.cu) adds the 4 cells around every
// logit. A CELL falls back to DIRECT evaluation (FFMA + MUFU per class-pixel against exact per-pixel maxima) when the
// chain could leave the normal range: a corner more than 100 log2-units below the cell maximum, or horizontal corner
// differences so large that R^(PXC/2) could overflow.
//
// Top-1: the label's class is the arg-max iff its exponential reaches the pixel's maximum exponential up to the chain's
// rounding (2^-19 relative): exact and near ties count as correct (torch.topk's choice among ties is unspecified).
// Bound (ncu, config 2): the L1 / shared-memory data pipe at 67 % (every 16-byte-per-lane shared access is 4 wavefronts:
// the private sums' read-modify-write per class and row is the bulk) and instruction issue at 57 %; HBM traffic is the
// label map. What the data pipe dictated: 16-byte label loads, a bank-conflict-free column order in the cell write-out.
// Algorithmic bytes per launch: 2*N*C*h*w*s + N*H*W*L.
#pragma once
#include "common.cuh"
#include "loss_upcell.cuh"   // RawLabel / decode_label / lg2 / pixel_weight / kLn2

namespace b200seg {

constexpr int kUpBatch = 4;               // classes in flight per thread in the class sweeps: their shared-memory loads are issued
                                          // together, the 4 dependent chains (LDS -> FFMA2 -> MUFU -> Horner) interleave
constexpr float kUpPadCorner = -150.f;    // pad classes (class count rounded up to the batch): 2^-150 flushes to an exact 0
__host__ __device__ constexpr int up_pad4(int c) { return (c + kUpBatch - 1) & ~(kUpBatch - 1); }

// n / d for 2 <= d <= 2^31 and n < 2^31 as __umulhi(n, mul) >> shr (round-up magic number; checked exhaustively over
// d < 3000 and at random by tools/probe/fastdiv_check.py)
struct UpFastDiv { unsigned mul, shr; };
static inline UpFastDiv up_fastdiv(unsigned d) {
  int s = 0;
  while (s < 31 && (1u << s) < d) ++s;
  UpFastDiv f;
  f.mul = (unsigned)((((unsigned long long)1 << (31 + s)) + d - 1) / d);
  f.shr = (unsigned)(s - 1);
  return f;
}

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
  float inv_sh, inv_sw;   // 1 / scale (0 when the scale is 0): first guess of a band start, corrected against the map itself
  int RG, logRG;      // threads per cell (row groups)
  long long cells;    // N * (h + 1) * (w + 1)
  UpFastDiv div_w1, div_h1;   // division by w + 1 / h + 1
  int ignore32;       // ignore_index as int32 (kNeverLabel if it does not fit: never matches)
  int acc_ignore32;   // accuracy's ignore_index, kNeverLabel when it has none
};

// ---- shared memory by 32-bit shared-window address. (With generic pointers into the dynamic segment the compiler re-derived
// the window base from SR_CgaCtaId at every use in the straight-line pixel code: 4 issue slots per access.)
__device__ __forceinline__ float4 lds4(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts4(unsigned a, const float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// predicated forms (v keeps its value / nothing is stored when flag == 0): straight-line code, no branch
__device__ __forceinline__ void lds4_if(unsigned a, int flag, float4& v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
               : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
               : "r"(a), "r"(flag));
}
__device__ __forceinline__ void sts4_if(unsigned a, int flag, const float4 v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t@p st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"r"(a), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "r"(flag)
               : "memory");
}
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

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
// first output position in [0, out] whose key is >= b (keys are non-decreasing in dst): a guess from the inverse scale,
// then corrected against the map itself
__device__ __forceinline__ int up_band_start(float scale, float inv_scale, int b, int in, int out, bool ac) {
  if (b <= 0) return 0;
  if (b > in || !(scale > 0.f)) return out;
  const float est = ac ? ((float)(b - 1) * inv_scale) : (((float)(b - 1) + 0.5f) * inv_scale - 0.5f);
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

// e[2k], e[2k+1] = E0 * R^(2k), E0 * R^(2k+1) as pixel pairs, product tree of depth <= 3
template <int PXC> __device__ __forceinline__ void up_chain2(float E0, float R, float2 (&e)[PXC / 2]) {
  static_assert(PXC == 4 || PXC == 8, "chunk width");
  const float2 R2 = __fmul2_rn(f2(R, R), f2(R, R));
  e[0] = f2(E0, E0 * R);
  e[1] = __fmul2_rn(e[0], R2);
  if constexpr (PXC == 8) {
    const float2 R4 = __fmul2_rn(R2, R2);
    e[2] = __fmul2_rn(e[0], R4);
    e[3] = __fmul2_rn(e[1], R4);
  }
}
// largest |D2 * sx| for which R^(PXC/2) stays a normal number
template <int PXC> constexpr float kUpChainLimit = PXC == 8 ? 30.f : 60.f;

template <int THR> constexpr size_t upgen_smem_bytes(int C, int logRG, bool grad) {
  return (size_t)up_pad4(C) * (THR >> logRG) * 16 + (grad ? (size_t)up_pad4(C) * THR * 16 : 0);
}

// ---- which cell this thread works on, and which of its rows
struct UpCellGeom {
  int n, b, r, rg, cell;
  bool ok, yreg;
  int Yr0, Yr1, X0, X1;
  float sx, lx0, ly_clamped;
};

// which cell (enough to request its corner logits) ...
template <int THR> __device__ __forceinline__ UpCellGeom up_cell_ids(const UpGenParams& p, int tid) {
  UpCellGeom g;
  const int RG = p.RG;
  const int cpc = THR >> p.logRG;
  g.rg = tid & (RG - 1);
  g.cell = tid >> p.logRG;
  const unsigned cid_raw = blockIdx.x * (unsigned)cpc + (unsigned)g.cell;
  g.ok = cid_raw < (unsigned)p.cells;
  const unsigned cid = g.ok ? cid_raw : (unsigned)p.cells - 1u;
  const unsigned t0 = __umulhi(cid, p.div_w1.mul) >> p.div_w1.shr;
  g.r = (int)(cid - t0 * (unsigned)(p.w + 1));
  g.n = (int)(__umulhi(t0, p.div_h1.mul) >> p.div_h1.shr);
  g.b = (int)(t0 - (unsigned)g.n * (unsigned)(p.h + 1));
  return g;
}
// ... and its extents / this thread's rows (computed while the corner loads are in flight)
template <int THR> __device__ __forceinline__ void up_cell_extents(const UpGenParams& p, UpCellGeom& g, int tid) {
  const int RG = p.RG;
  const bool ac = p.ac != 0;
  int Yb0, Yb1;
  if (RG == 4) {
    // the 4 threads of a cell compute one extent each
    const bool isx = g.rg >= 2;
    const int v = up_band_start(isx ? p.sw : p.sh, isx ? p.inv_sw : p.inv_sh, (isx ? g.r : g.b) + (g.rg & 1), isx ? p.w : p.h,
                                isx ? p.W : p.H, ac);
    const int l0 = (tid & 31) & ~3;
    Yb0 = __shfl_sync(0xffffffffu, v, l0);
    Yb1 = __shfl_sync(0xffffffffu, v, l0 + 1);
    g.X0 = __shfl_sync(0xffffffffu, v, l0 + 2);
    g.X1 = __shfl_sync(0xffffffffu, v, l0 + 3);
  } else {
    Yb0 = up_band_start(p.sh, p.inv_sh, g.b, p.h, p.H, ac);
    Yb1 = up_band_start(p.sh, p.inv_sh, g.b + 1, p.h, p.H, ac);
    g.X0 = up_band_start(p.sw, p.inv_sw, g.r, p.w, p.W, ac);
    g.X1 = up_band_start(p.sw, p.inv_sw, g.r + 1, p.w, p.W, ac);
  }
  const int rows_per = (Yb1 - Yb0 + RG - 1) >> p.logRG;
  g.Yr0 = Yb0 + g.rg * rows_per;
  g.Yr1 = g.ok ? min(Yb1, g.Yr0 + rows_per) : g.Yr0;
  // regular run: lambda advances by the scale. Clamped runs / bands read one tap only: the weight is put on the tap the
  // combine step reads for that cell — the HIGH tap of band 0 / run 0 (row 0 / column 0), the LOW tap of band h / run w.
  const bool xreg = (g.r > 0 && g.r < p.w);
  g.sx = xreg ? p.sw : 0.f;
  g.lx0 = xreg ? (g.X0 < g.X1 ? up_lambda(p.sw, g.X0, g.r - 1, ac) : 0.f) : (g.r == 0 ? 1.f : 0.f);
  g.yreg = (g.b > 0 && g.b < p.h);
  g.ly_clamped = (g.b == 0) ? 1.f : 0.f;
}

// ---- the cell's 4 corner logits of the classes [c0, c0 + ct) into CORN[c][cell] = (a, b, da, db) scaled by log2 e and
// offset by the maximum; the RG threads of the cell split the classes (5 classes = 20 loads in flight per thread: all of a
// 19- or 20-class cell's share in one round trip, requested BEFORE the cell's extents are worked out). Rows
// [ct, up_pad4(ct)) become pad classes. Returns, reduced over the cell: the maximum M, the minimum mn, the largest
// |horizontal difference| hd (natural units).
constexpr int kUpLoadBatch = 5;
template <typename T> struct UpCornerLoad {
  const T* lg;
  unsigned o00, o01, o10, o11, plane;
  float v[kUpLoadBatch][4];
  __device__ __forceinline__ void init(const UpGenParams& p, const UpCellGeom& g, int c0) {
    plane = (unsigned)(p.h * p.w);
    lg = reinterpret_cast<const T*>(p.logits);
    const int ya = g.b - 1 < 0 ? 0 : g.b - 1, yb = g.b > p.h - 1 ? p.h - 1 : g.b;
    const int xa = g.r - 1 < 0 ? 0 : g.r - 1, xb = g.r > p.w - 1 ? p.w - 1 : g.r;
    // element indices fit 32 bits (checked on the host): one IMAD + one wide add per load
    const unsigned base = ((unsigned)g.n * (unsigned)p.C + (unsigned)c0) * plane;
    o00 = base + (unsigned)(ya * p.w + xa); o01 = base + (unsigned)(ya * p.w + xb);
    o10 = base + (unsigned)(yb * p.w + xa); o11 = base + (unsigned)(yb * p.w + xb);
  }
  __device__ __forceinline__ void issue(int cb, int ct, int RG) {
#pragma unroll
    for (int u = 0; u < kUpLoadBatch; ++u) {
      const int c = cb + u * RG;
      const unsigned co = (unsigned)(c < ct ? c : 0) * plane;       // idle slots re-read class 0 (always in range)
      v[u][0] = to_float<T>(__ldg(lg + (o00 + co)));
      v[u][1] = to_float<T>(__ldg(lg + (o01 + co)));
      v[u][2] = to_float<T>(__ldg(lg + (o10 + co)));
      v[u][3] = to_float<T>(__ldg(lg + (o11 + co)));
    }
  }
  // MODE 0: fold the batch into the running max / min / horizontal difference.
  // MODE 1: write the batch to CORN raw (a later pass normalises in place).   MODE 2: write it normalised.
  template <int MODE>
  __device__ __forceinline__ void consume(int cb, int ct, int RG, unsigned corn_cell, unsigned corn_stride, float& M, float& mn,
                                          float& hd, float nM2) {
#pragma unroll
    for (int u = 0; u < kUpLoadBatch; ++u) {
      const int c = cb + u * RG;
      if (c < ct) {
        if constexpr (MODE == 0) {
          M = fmaxf(M, fmaxf(fmaxf(v[u][0], v[u][1]), fmaxf(v[u][2], v[u][3])));
          mn = fminf(mn, fminf(fminf(v[u][0], v[u][1]), fminf(v[u][2], v[u][3])));
          hd = fmaxf(hd, fmaxf(fabsf(v[u][1] - v[u][0]), fabsf(v[u][3] - v[u][2])));
        } else if constexpr (MODE == 1) {
          sts4(corn_cell + (unsigned)c * corn_stride, make_float4(v[u][0], v[u][1], v[u][2], v[u][3]));
        } else {
          sts4(corn_cell + (unsigned)c * corn_stride,
               make_float4(fmaf(v[u][0], kLog2e, nM2), fmaf(v[u][1], kLog2e, nM2), (v[u][2] - v[u][0]) * kLog2e,
                           (v[u][3] - v[u][1]) * kLog2e));
        }
      }
    }
  }
};

// geometry + corner logits of a thread's cell (both kernels): ids, first batch of corner loads, extents, the rest
template <typename T, int THR>
__device__ __forceinline__ UpCellGeom up_cell_setup(const UpGenParams& p, int tid, int c0, int ct, unsigned sm0, unsigned corn_stride,
                                                    float& M, float& mn, float& hd) {
  UpCellGeom g = up_cell_ids<THR>(p, tid);
  const int RG = p.RG;
  const unsigned corn_cell = sm0 + (unsigned)g.cell * 16u;
  UpCornerLoad<T> cl;
  cl.init(p, g, c0);
  cl.issue(g.rg, ct, RG);
  up_cell_extents<THR>(p, g, tid);
  M = -3.0e38f;
  mn = 3.0e38f;
  hd = 0.f;
  const bool one_batch = ct <= kUpLoadBatch * RG;      // the usual case: the thread's share stays in registers
  cl.template consume<0>(g.rg, ct, RG, corn_cell, corn_stride, M, mn, hd, 0.f);
  if (!one_batch) {
    cl.template consume<1>(g.rg, ct, RG, corn_cell, corn_stride, M, mn, hd, 0.f);
    for (int cb = g.rg + kUpLoadBatch * RG; cb < ct; cb += kUpLoadBatch * RG) {
      cl.issue(cb, ct, RG);
      cl.template consume<0>(cb, ct, RG, corn_cell, corn_stride, M, mn, hd, 0.f);
      cl.template consume<1>(cb, ct, RG, corn_cell, corn_stride, M, mn, hd, 0.f);
    }
  }
  for (int off = 1; off < RG; off <<= 1) {
    M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, off));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    hd = fmaxf(hd, __shfl_xor_sync(0xffffffffu, hd, off));
  }
  const float nM2 = -M * kLog2e;
  if (one_batch) {
    cl.template consume<2>(g.rg, ct, RG, corn_cell, corn_stride, M, mn, hd, nM2);
  } else {
    for (int c = g.rg; c < ct; c += RG) {
      const unsigned a = corn_cell + (unsigned)c * corn_stride;
      const float4 v = lds4(a);
      sts4(a, make_float4(fmaf(v.x, kLog2e, nM2), fmaf(v.y, kLog2e, nM2), (v.z - v.x) * kLog2e, (v.w - v.y) * kLog2e));
    }
  }
  if (g.rg == 0) {
    for (int c = ct; c < up_pad4(ct); ++c) sts4(corn_cell + (unsigned)c * corn_stride, make_float4(kUpPadCorner, kUpPadCorner, 0.f, 0.f));
  }
  return g;
}

// ---- backward sweep over the classes [0, cp4) of one chunk (cp4 a multiple of kUpBatch; pad classes add exact zeros):
// sum_j a_j E0 R^j and sum_j a_j lambda_j E0 R^j are two polynomials in R (coefficients >= 0, no cancellation) evaluated
// together by one packed Horner recurrence, then folded into the class's sums in the thread's private column.
template <int PXC, int THR>
__device__ __forceinline__ void up_bwd_sweep(int cp4, unsigned corn, unsigned corn_stride, unsigned oh_col, float ly, float lam0,
                                             float sx, const float2 (&ab)[PXC]) {
  const float2 ly2 = f2(ly, ly), ly02 = f2(1.f - ly, 1.f - ly);
  // the corner logits of the NEXT batch are requested before the current batch's arithmetic (register double buffer: the
  // shared-memory latency was the top stall of the single-buffered loop); unrolled by 2 so that the rotation is a renaming
  float4 q[kUpBatch];
#pragma unroll
  for (int u = 0; u < kUpBatch; ++u) q[u] = lds4(corn + (unsigned)u * corn_stride);
#pragma unroll 2
  for (int c = 0; c < cp4; c += kUpBatch) {
    float4 qn[kUpBatch], o[kUpBatch];
    const int cn = c + kUpBatch < cp4 ? c + kUpBatch : c;
#pragma unroll
    for (int u = 0; u < kUpBatch; ++u) o[u] = lds4(oh_col + (unsigned)(c + u) * (unsigned)(THR * 16));
#pragma unroll
    for (int u = 0; u < kUpBatch; ++u) qn[u] = lds4(corn + (unsigned)(cn + u) * corn_stride);
#pragma unroll
    for (int u = 0; u < kUpBatch; ++u) {
      const float2 LR = __ffma2_rn(ly2, f2(q[u].z, q[u].w), f2(q[u].x, q[u].y));
      const float D2 = LR.y - LR.x;
      const float E0 = ex2(fmaf(lam0, D2, LR.x)), R = ex2(D2 * sx);
      const float2 RR = f2(R, R);
      float2 P = ab[PXC - 1];
#pragma unroll
      for (int j = PXC - 2; j >= 0; --j) P = __ffma2_rn(P, RR, ab[j]);
      const float2 G = __fmul2_rn(P, f2(E0, E0));      // (sum softmax weight, sum softmax weight * lambda) of the row
      const float2 o1 = __ffma2_rn(ly02, G, f2(o[u].x, o[u].y)), o2 = __ffma2_rn(ly2, G, f2(o[u].z, o[u].w));
      o[u] = make_float4(o1.x, o1.y, o2.x, o2.y);
    }
#pragma unroll
    for (int u = 0; u < kUpBatch; ++u) sts4(oh_col + (unsigned)(c + u) * (unsigned)(THR * 16), o[u]);
#pragma unroll
    for (int u = 0; u < kUpBatch; ++u) q[u] = qn[u];
  }
}

// ---- the RG private columns of a cell are added in a fixed order: thread rg takes the classes c = rg (mod RG) and reads the
// RG columns straight from shared memory (no shuffle tree), then writes those classes of the cell to PB. The columns hold
// (S, SL) pairs per tap row — sums of (softmax - onehot) and of (softmax - onehot) * lambda — the corner sums are (S - SL, SL).
template <int RGC, int THR>
__device__ __forceinline__ void up_write_cell(const UpGenParams& p, const UpCellGeom& g, int tid, unsigned oh_base, int c0, int ct) {
  const unsigned col0 = oh_base + (unsigned)(tid - g.rg) * 16u;
  float4* pbc = reinterpret_cast<float4*>(p.pb) + (((size_t)g.n * p.C + c0) * (p.h + 1) + g.b) * (p.w + 1) + g.r;
  const size_t cstride = (size_t)(p.h + 1) * (p.w + 1);
  for (int c = g.rg; c < ct; c += 2 * RGC) {
    const int c2 = c + RGC;
    const bool two = c2 < ct;
    const unsigned oa = col0 + (unsigned)c * (unsigned)(THR * 16), ob = col0 + (unsigned)(two ? c2 : c) * (unsigned)(THR * 16);
    // the RGC threads of a cell read different classes (512 bytes apart: the same banks), so thread rg starts at column rg:
    // conflict-free, and still a fixed summation order per class
    float4 t[RGC], t2[RGC];
#pragma unroll
    for (int k = 0; k < RGC; ++k) t[k] = lds4(oa + (unsigned)((k + g.rg) & (RGC - 1)) * 16u);
#pragma unroll
    for (int k = 0; k < RGC; ++k) t2[k] = lds4(ob + (unsigned)((k + g.rg) & (RGC - 1)) * 16u);
    float4 o = t[0], o2 = t2[0];
#pragma unroll
    for (int k = 1; k < RGC; ++k) {
      o.x += t[k].x; o.y += t[k].y; o.z += t[k].z; o.w += t[k].w;
      o2.x += t2[k].x; o2.y += t2[k].y; o2.z += t2[k].z; o2.w += t2[k].w;
    }
    pbc[(size_t)c * cstride] = make_float4(o.x - o.y, o.y, o.z - o.w, o.w);
    if (two) pbc[(size_t)c2 * cstride] = make_float4(o2.x - o2.y, o2.y, o2.z - o2.w, o2.w);
  }
}

// out-of-range labels of a chunk (neither a class nor ignore_index): the rare path behind the per-chunk `any bad` flag
template <int LK> static __device__ __noinline__ int up_count_bad(const char* labimg, int dt, unsigned px0, int npx, int C,
                                                                  int ignore32) {
  int n = 0;
  for (int j = 0; j < npx; ++j) {
    const int yy = decode_label<LK>(load_raw_label<LK>(labimg, dt, px0 + (unsigned)j), dt);
    n += (yy != ignore32 && (unsigned)yy >= (unsigned)C);
  }
  return n;
}

// ---- the PXC labels of a chunk. The kernel is bound by the L1 / shared-memory data pipe, and 8-byte label loads scattered
// over the rows and cells of a warp cost it ~23 wavefronts per request: a full chunk is fetched with 16-byte loads (int64:
// two labels, uint8: four) whenever its first label is suitably aligned, a quarter / half of the requests.
template <int PXC, int LK>
__device__ __forceinline__ void up_load_labels(const char* labimg, int dt, unsigned px0, int npx, RawLabel (&raw)[PXC]) {
  if constexpr (LK == 0) {
    const char* a = labimg + (size_t)px0 * 8;
    if (npx == PXC && (reinterpret_cast<size_t>(a) & 15) == 0) {
#pragma unroll
      for (int k = 0; k < PXC / 2; ++k) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(a) + k);
        raw[2 * k].lo = v.x; raw[2 * k].hi = v.y;
        raw[2 * k + 1].lo = v.z; raw[2 * k + 1].hi = v.w;
      }
      return;
    }
  }
  if constexpr (LK == 1) {
    const char* a = labimg + px0;
    if (npx == PXC && (reinterpret_cast<size_t>(a) & 3) == 0) {
#pragma unroll
      for (int k = 0; k < PXC / 4; ++k) {
        const unsigned v = __ldg(reinterpret_cast<const unsigned*>(a) + k);
#pragma unroll
        for (int i = 0; i < 4; ++i) { raw[4 * k + i].lo = (v >> (8 * i)) & 0xffu; raw[4 * k + i].hi = 0; }
      }
      return;
    }
  }
#pragma unroll
  for (int j = 0; j < PXC; ++j) raw[j] = load_raw_label<LK>(labimg, dt, px0 + (unsigned)min(j, npx - 1));
}

struct UpAcc {
  float loss;
  int n_valid, n_correct, n_acc, n_bad;
};

// ---- per-pixel pass of one chunk: label, its interpolated logit, loss, top-1, the softmax weights (a, a lambda) of the
// backward sweep, then the one-hot terms. Straight-line code (the weighted case arrives through `wts`): the first loop
// holds no store, so the 8 pixels' dependent chains (LDS -> FFMA2 -> FFMA -> MUFU) overlap; the one-hot terms of a RUN of
// equal labels are added up in registers and folded into the label class's private sums once per run — its column entry
// is requested when the run starts and written when it ends (predicated, no branch), so no shared-memory round trip sits
// between consecutive pixels. OFF: direct evaluation (per-pixel offsets, clamped lambdas).
template <int PXC, bool GRAD, int LK, bool OFF, int THR>
__device__ __forceinline__ bool up_pixel_pass(const UpGenParams& p, const RawLabel (&raw)[PXC], const float (&wts)[PXC], int npx,
                                              int dt, int C, unsigned corn, unsigned corn_stride, unsigned oh_col, float ly,
                                              float lam0, float sx, const float (&s)[PXC], const float (&m)[PXC],
                                              const float (&moff)[PXC], float2 (&ab)[PXC], UpAcc& acc, float* lse_row,
                                              float M2cell) {
  const float2 ly2 = f2(ly, ly);
  const float jcap = (float)(npx - 1);
  const int ign32 = p.ignore32, accign32 = p.acc_ignore32;
  bool anybad = false;
  int yc[PXC];
  float2 wu[PXC];
#pragma unroll
  for (int j = 0; j < PXC; ++j) {
    const bool ok = j < npx;
    int yy = decode_label<LK>(raw[j], dt);
    yy = ok ? yy : ign32;
    const bool ign = (yy == ign32);
    const bool inr = (unsigned)yy < (unsigned)C;
    const bool use = inr && !ign;
    // a pixel without a class (ignored, masked, out of range) stays in the run of its left neighbour with weight 0
    yc[j] = use ? yy : (j ? yc[j - 1] : 0);
    const bool accp = (yy != accign32);
    acc.n_valid += !ign;
    acc.n_acc += accp;
    anybad = anybad || (!ign && !inr);
    const float wt = use ? wts[j] : 0.f;
    const float lamj = OFF ? fmaf(fminf((float)j, jcap), sx, lam0) : fmaf((float)j, sx, lam0);
    // the label's interpolated logit and exponential: the operations of the class sweep on the label's corners
    const float4 q = lds4(corn + (unsigned)yc[j] * corn_stride);
    const float2 LR = __ffma2_rn(ly2, f2(q.z, q.w), f2(q.x, q.y));
    const float D2 = LR.y - LR.x;
    const float zy2 = fmaf(lamj, D2, LR.x);
    const float sj = ok ? s[j] : 1.f;   // the masked tail of a partial chunk may hold 0 / inf
    // the label is the arg-max iff its exponential reaches the pixel's maximum exponential; the maximum comes out of
    // the product chain (relative error <= 2^-20: four ex2.approx factors), so the test allows 2^-19 — an exact or
    // near tie with another class counts as correct for the label (torch.topk's choice among ties is unspecified)
    const float ey = ex2(OFF ? zy2 - moff[j] : zy2);
    const float lse2_rel = OFF ? moff[j] + lg2(sj) : lg2(sj);
    acc.loss = fmaf(wt, lse2_rel - zy2, acc.loss);
    acc.n_correct += (use && accp && ey >= m[j] * 0.99999809265f);
    if constexpr (!GRAD) {
      if (lse_row && ok) lse_row[j] = M2cell + lse2_rel;
    }
    if constexpr (GRAD) {
      const float a = wt * fast_rcp(sj);
      ab[j] = f2(a, a * lamj);
      wu[j] = f2(wt, wt * lamj);
    }
  }
  // the masked pixels carry ignore_index: they were counted above when accuracy ignores a different value
  acc.n_acc -= (ign32 != accign32) ? (PXC - npx) : 0;
  if constexpr (GRAD) {
    // one-hot terms, run by run (sums are kept as sums of softmax - onehot and of (softmax - onehot) lambda)
    const float2 nly0 = f2(ly - 1.f, ly - 1.f), nly = f2(-ly, -ly);
    float2 run = wu[0];
    unsigned ra = oh_col + (unsigned)yc[0] * (unsigned)(THR * 16);
    float4 o = lds4(ra);
#pragma unroll
    for (int j = 1; j < PXC; ++j) {
      const int chg = (yc[j] != yc[j - 1]);
      const float2 t1 = __ffma2_rn(nly0, run, f2(o.x, o.y)), t2 = __ffma2_rn(nly, run, f2(o.z, o.w));
      sts4_if(ra, chg, make_float4(t1.x, t1.y, t2.x, t2.y));      // the run ended: its class's sums go back
      ra = oh_col + (unsigned)yc[j] * (unsigned)(THR * 16);
      lds4_if(ra, chg, o);                                         // ... and the new run's class is requested
      const float keep = chg ? 0.f : 1.f;
      run = __ffma2_rn(run, f2(keep, keep), wu[j]);
    }
    const float2 t1 = __ffma2_rn(nly0, run, f2(o.x, o.y)), t2 = __ffma2_rn(nly, run, f2(o.z, o.w));
    sts4(ra, make_float4(t1.x, t1.y, t2.x, t2.y));
  }
  return anybad;
}

template <typename T, int PXC, bool GRAD, int LK, int THR>
__global__ void __launch_bounds__(THR) up_gen_kernel(const UpGenParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();   // finalize may be scheduled as soon as SM resources free up; it waits for this grid itself
  const unsigned sm0 = (unsigned)__cvta_generic_to_shared(smem_raw);
  const int tid = threadIdx.x;
  const int C = p.C;
  const int Cp4 = up_pad4(C);
  const int RG = p.RG;
  const int cpc = THR >> p.logRG;                                  // cells per CTA
  const unsigned corn_stride = (unsigned)cpc * 16u;               // CORN [Cp4][cpc] float4: (a, b, da, db) per class of the cell
  const unsigned oh_base = sm0 + (unsigned)Cp4 * corn_stride;     // OH   [Cp4][THR] float4: private sums (GRAD)
  float M, mn, hd;
  const UpCellGeom g = up_cell_setup<T, THR>(p, tid, 0, C, sm0, corn_stride, M, mn, hd);
  const unsigned corn = sm0 + (unsigned)g.cell * 16u;
  const unsigned oh_col = oh_base + (unsigned)tid * 16u;
  if constexpr (GRAD) {
    for (int c = 0; c < Cp4; ++c) sts4(oh_col + (unsigned)c * (unsigned)(THR * 16), make_float4(0.f, 0.f, 0.f, 0.f));
  }
  __syncwarp();
  const float M2cell = M * kLog2e;
  const float sx = g.sx;
  // every tap within 100 log2-units of the cell maximum (exponentials and the Horner partial sums <= weight / E0 stay
  // normal), and R^(PXC/2) a normal number for every class and row
  const bool cell_fast = ((M - mn) * kLog2e <= 100.f) && (hd * kLog2e * sx <= kUpChainLimit<PXC>);

  UpAcc acc;
  acc.loss = 0.f;
  acc.n_valid = acc.n_correct = acc.n_acc = acc.n_bad = 0;
  const bool ac = p.ac != 0;
  const int dt = p.label_dtype;
  const int lb = LK == 0 ? 8 : (LK == 1 ? 1 : p.label_bytes);
  const size_t img_px = (size_t)g.n * p.H * p.W;
  const char* labimg = reinterpret_cast<const char*>(p.labels) + img_px * lb;

#pragma unroll 1
  for (int Y = g.Yr0; Y < g.Yr1; ++Y) {
    const float ly = g.yreg ? up_lambda(p.sh, Y, g.b - 1, ac) : g.ly_clamped;
    const float2 ly2 = f2(ly, ly);
    const unsigned roff = (unsigned)Y * (unsigned)p.W;
#pragma unroll 1
    for (int Xc = g.X0; Xc < g.X1; Xc += PXC) {
      const int npx = min(PXC, g.X1 - Xc);
      // ---- labels of the chunk, issued before the class sweep (consumed after it)
      RawLabel raw[PXC];
      up_load_labels<PXC, LK>(labimg, dt, roff + (unsigned)Xc, npx, raw);
      const float lam0 = fmaf((float)(Xc - g.X0), sx, g.lx0);
      float wts[PXC];
#pragma unroll
      for (int j = 0; j < PXC; ++j) wts[j] = 1.f;
      if (p.has_w) {
#pragma unroll
        for (int j = 0; j < PXC; ++j) {
          const int yy = decode_label<LK>(raw[j], dt);
          const bool use = (j < npx) && yy != p.ignore32 && (unsigned)yy < (unsigned)C;
          wts[j] = pixel_weight(p.cw, p.pw, use, use ? yy : 0, img_px + (size_t)(roff + (unsigned)(Xc + min(j, npx - 1))));
        }
      }
      float* lse_row = (!GRAD && p.lse2) ? p.lse2 + img_px + (size_t)(roff + (unsigned)Xc) : nullptr;
      float2 ab[PXC];
      bool anybad;
      if (cell_fast) {
        // ---- forward sweep, geometric chain, kUpBatch classes per step (the pad classes add exact zeros)
        float2 s2[PXC / 2];
        float m[PXC];
#pragma unroll
        for (int k = 0; k < PXC / 2; ++k) s2[k] = f2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < PXC; ++j) m[j] = 0.f;
        float4 q[kUpBatch];
#pragma unroll
        for (int u = 0; u < kUpBatch; ++u) q[u] = lds4(corn + (unsigned)u * corn_stride);
#pragma unroll 2
        for (int c = 0; c < Cp4; c += kUpBatch) {
          float4 qn[kUpBatch];      // next batch's corners, requested before this batch's arithmetic
          const int cn = c + kUpBatch < Cp4 ? c + kUpBatch : c;
#pragma unroll
          for (int u = 0; u < kUpBatch; ++u) qn[u] = lds4(corn + (unsigned)(cn + u) * corn_stride);
          float2 e[kUpBatch][PXC / 2];
#pragma unroll
          for (int u = 0; u < kUpBatch; ++u) {
            const float2 LR = __ffma2_rn(ly2, f2(q[u].z, q[u].w), f2(q[u].x, q[u].y));
            const float D2 = LR.y - LR.x;
            up_chain2<PXC>(ex2(fmaf(lam0, D2, LR.x)), ex2(D2 * sx), e[u]);
#pragma unroll
            for (int k = 0; k < PXC / 2; ++k) s2[k] = __fadd2_rn(s2[k], e[u][k]);
          }
#pragma unroll
          for (int u = 0; u < kUpBatch; u += 2) {
#pragma unroll
            for (int k = 0; k < PXC / 2; ++k) {
              m[2 * k] = fmaxf(m[2 * k], fmaxf(e[u][k].x, e[u + 1][k].x));
              m[2 * k + 1] = fmaxf(m[2 * k + 1], fmaxf(e[u][k].y, e[u + 1][k].y));
            }
          }
#pragma unroll
          for (int u = 0; u < kUpBatch; ++u) q[u] = qn[u];
        }
        float s[PXC], moff[PXC];
#pragma unroll
        for (int k = 0; k < PXC / 2; ++k) { s[2 * k] = s2[k].x; s[2 * k + 1] = s2[k].y; }
#pragma unroll
        for (int j = 0; j < PXC; ++j) moff[j] = 0.f;
        anybad = up_pixel_pass<PXC, GRAD, LK, false, THR>(p, raw, wts, npx, dt, C, corn, corn_stride, oh_col, ly, lam0, sx, s, m, moff,
                                                          ab, acc, lse_row, M2cell);
        if constexpr (GRAD) up_bwd_sweep<PXC, THR>(Cp4, corn, corn_stride, oh_col, ly, lam0, sx, ab);
      } else {
        // ---- direct evaluation against exact per-pixel maxima (lambda of the masked tail pixels = the last valid one)
        float s[PXC], m[PXC], moff[PXC], lamc[PXC];
#pragma unroll
        for (int j = 0; j < PXC; ++j) {
          s[j] = 0.f; m[j] = 0.f; moff[j] = -3.0e38f;
          lamc[j] = fmaf((float)min(j, npx - 1), sx, lam0);
        }
        for (int c = 0; c < C; ++c) {
          const float4 q = lds4(corn + (unsigned)c * corn_stride);
          const float L2 = fmaf(ly, q.z, q.x), D2 = fmaf(ly, q.w, q.y) - L2;
#pragma unroll
          for (int j = 0; j < PXC; ++j) moff[j] = fmaxf(moff[j], fmaf(lamc[j], D2, L2));
        }
        for (int c = 0; c < C; ++c) {
          const float4 q = lds4(corn + (unsigned)c * corn_stride);
          const float L2 = fmaf(ly, q.z, q.x), D2 = fmaf(ly, q.w, q.y) - L2;
#pragma unroll
          for (int j = 0; j < PXC; ++j) {
            const float ev = ex2(fmaf(lamc[j], D2, L2) - moff[j]);
            s[j] += ev;
            m[j] = fmaxf(m[j], ev);
          }
        }
        anybad = up_pixel_pass<PXC, GRAD, LK, true, THR>(p, raw, wts, npx, dt, C, corn, corn_stride, oh_col, ly, lam0, sx, s, m, moff,
                                                         ab, acc, lse_row, M2cell);
        if constexpr (GRAD) {
          for (int c = 0; c < C; ++c) {
            const float4 q = lds4(corn + (unsigned)c * corn_stride);
            const float L2 = fmaf(ly, q.z, q.x), D2 = fmaf(ly, q.w, q.y) - L2;
            float gs = 0.f, gb = 0.f;
#pragma unroll
            for (int j = 0; j < PXC; ++j) {
              const float ev = ex2(fmaf(lamc[j], D2, L2) - moff[j]);
              gs = fmaf(ev, ab[j].x, gs);
              gb = fmaf(ev, ab[j].y, gb);
            }
            const unsigned oa = oh_col + (unsigned)c * (unsigned)(THR * 16);
            float4 o = lds4(oa);
            o.x = fmaf(1.f - ly, gs, o.x);
            o.y = fmaf(1.f - ly, gb, o.y);
            o.z = fmaf(ly, gs, o.z);
            o.w = fmaf(ly, gb, o.w);
            sts4(oa, o);
          }
        }
      }
      if (anybad) acc.n_bad += up_count_bad<LK>(labimg, dt, roff + (unsigned)Xc, npx, C, p.ignore32);
    }
  }

  if constexpr (GRAD) {
    __syncwarp();
    if (g.ok) {
      if (RG == 4) up_write_cell<4, THR>(p, g, tid, oh_base, 0, C);
      else if (RG == 2) up_write_cell<2, THR>(p, g, tid, oh_base, 0, C);
      else up_write_cell<1, THR>(p, g, tid, oh_base, 0, C);
    }
  }
  static_assert(THR == 32, "one warp per CTA (warp_flush_stats)");
  warp_flush_stats(acc.loss * kLn2, acc.n_valid, acc.n_correct, acc.n_bad, acc.n_acc, p.stats);
}

// ------------------------------------------------------------------------------------------------ class-tiled backward
// C > 32: a thread's private corner sums for ALL classes do not fit shared memory (16 C bytes per thread), so the plan is
// split: one forward-only launch of up_gen_kernel over all classes (loss, accuracy, and the per-pixel log2-sum-exp into
// `lse2`), then this kernel once per tile of kUpTile classes (grid.y): soft-max probabilities of the tile's classes from
// the saved lse2 — p = 2^(z2 - lse2) — the same packed Horner recurrence / corner-sum machinery, the tile's slice of PB.
// Deterministic.
constexpr int kUpTile = 32;

template <typename T, int PXC, int LK, int THR>
__global__ void __launch_bounds__(THR) up_gen_bwd_tile_kernel(const UpGenParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned sm0 = (unsigned)__cvta_generic_to_shared(smem_raw);
  const int tid = threadIdx.x;
  const int C = p.C;
  const int RG = p.RG;
  const int cpc = THR >> p.logRG;
  const unsigned corn_stride = (unsigned)cpc * 16u;               // CORN [kUpTile][cpc]
  const unsigned oh_base = sm0 + (unsigned)kUpTile * corn_stride; // OH   [kUpTile][THR]
  const int c0 = blockIdx.y * kUpTile;
  const int ct = min(kUpTile, C - c0);
  float M, mn, hd;
  const UpCellGeom g = up_cell_setup<T, THR>(p, tid, c0, ct, sm0, corn_stride, M, mn, hd);
  const unsigned corn = sm0 + (unsigned)g.cell * 16u;
  const unsigned oh_col = oh_base + (unsigned)tid * 16u;
  const int ct4 = up_pad4(ct);
  for (int c = 0; c < ct4; ++c) sts4(oh_col + (unsigned)c * (unsigned)(THR * 16), make_float4(0.f, 0.f, 0.f, 0.f));
  __syncwarp();
  const float M2t = M * kLog2e;
  const float sx = g.sx;
  // the chain, the Horner partial sums and the per-pixel factor 2^(M2t - lse2) <= 2^(M2t - min) stay in the normal range
  const bool chain_ok = ((M - mn) * kLog2e <= 100.f) && (hd * kLog2e * sx <= 100.f);

  const bool ac = p.ac != 0;
  const int dt = p.label_dtype;
  const int lb = LK == 0 ? 8 : (LK == 1 ? 1 : p.label_bytes);
  const size_t img_px = (size_t)g.n * p.H * p.W;
  const char* labimg = reinterpret_cast<const char*>(p.labels) + img_px * lb;
  const float* lseimg = p.lse2 + img_px;

#pragma unroll 1
  for (int Y = g.Yr0; Y < g.Yr1; ++Y) {
    const float ly = g.yreg ? up_lambda(p.sh, Y, g.b - 1, ac) : g.ly_clamped;
    const float2 nly0 = f2(ly - 1.f, ly - 1.f), nly = f2(-ly, -ly);
    const unsigned roff = (unsigned)Y * (unsigned)p.W;
#pragma unroll 1
    for (int Xc = g.X0; Xc < g.X1; Xc += PXC) {
      const int npx = min(PXC, g.X1 - Xc);
      RawLabel raw[PXC];
      float off2[PXC];
#pragma unroll
      for (int j = 0; j < PXC; ++j) {
        const unsigned px = roff + (unsigned)(Xc + min(j, npx - 1));
        raw[j] = load_raw_label<LK>(labimg, dt, px);
        off2[j] = M2t - __ldg(lseimg + px);           // z2_rel + off2 = z2 - lse2 <= 0
      }
      const float lam0 = fmaf((float)(Xc - g.X0), sx, g.lx0);
      float2 ab[PXC];
      float wts[PXC];
#pragma unroll
      for (int j = 0; j < PXC; ++j) {
        const bool ok = j < npx;
        const int ydec = decode_label<LK>(raw[j], dt);
        const int yy = ok ? ydec : p.ignore32;
        const bool use = (yy != p.ignore32) && ((unsigned)yy < (unsigned)C);
        const int yc = use ? yy : 0;
        float wt = use ? 1.f : 0.f;
        if (p.has_w) wt = pixel_weight(p.cw, p.pw, use, yc, img_px + (size_t)(roff + (unsigned)(Xc + min(j, npx - 1))));
        const float lamj = fmaf((float)(chain_ok ? j : min(j, npx - 1)), sx, lam0);
        wts[j] = wt;
        const float a = chain_ok ? wt * ex2(off2[j]) : wt;   // direct evaluation folds off2 into the exponent instead
        ab[j] = f2(a, a * lamj);
        const int yl = yc - c0;
        if (use && (unsigned)yl < (unsigned)ct) {     // one-hot term, for labels of this tile
          const float2 wu = f2(wt, wt * lamj);
          const unsigned oa = oh_col + (unsigned)yl * (unsigned)(THR * 16);
          const float4 o = lds4(oa);
          const float2 o1 = __ffma2_rn(nly0, wu, f2(o.x, o.y)), o2 = __ffma2_rn(nly, wu, f2(o.z, o.w));
          sts4(oa, make_float4(o1.x, o1.y, o2.x, o2.y));
        }
      }
      if (chain_ok) {
        up_bwd_sweep<PXC, THR>(ct4, corn, corn_stride, oh_col, ly, lam0, sx, ab);
      } else {
        for (int c = 0; c < ct; ++c) {
          const float4 q = lds4(corn + (unsigned)c * corn_stride);
          const float L2 = fmaf(ly, q.z, q.x), D2 = fmaf(ly, q.w, q.y) - L2;
          float gs = 0.f, gb = 0.f;
#pragma unroll
          for (int j = 0; j < PXC; ++j) {
            const float ev = wts[j] != 0.f ? ex2(fmaf(fmaf((float)min(j, npx - 1), sx, lam0), D2, L2) + off2[j]) : 0.f;
            gs = fmaf(ev, ab[j].x, gs);
            gb = fmaf(ev, ab[j].y, gb);
          }
          const unsigned oa = oh_col + (unsigned)c * (unsigned)(THR * 16);
          float4 o = lds4(oa);
          o.x = fmaf(1.f - ly, gs, o.x);
          o.y = fmaf(1.f - ly, gb, o.y);
          o.z = fmaf(ly, gs, o.z);
          o.w = fmaf(ly, gb, o.w);
          sts4(oa, o);
        }
      }
    }
  }
  __syncwarp();
  if (g.ok) {
    if (RG == 4) up_write_cell<4, THR>(p, g, tid, oh_base, c0, ct);
    else if (RG == 2) up_write_cell<2, THR>(p, g, tid, oh_base, c0, ct);
    else up_write_cell<1, THR>(p, g, tid, oh_base, c0, ct);
  }
}

// resident CTAs per SM from the shared-memory footprint (228 KB per SM, 1 KB reserved per CTA) and the register file
template <int THR> static int upgen_resident_warps(int C, int logRG, bool grad) {
  const size_t smem = upgen_smem_bytes<THR>(C, logRG, grad) + 1024;
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
  p.inv_sh = p.sh > 0.f ? 1.f / p.sh : 0.f;
  p.inv_sw = p.sw > 0.f ? 1.f / p.sw : 0.f;
  p.RG = 1; p.logRG = 0;
  p.cells = (long long)f->N * (f->h + 1) * (f->w + 1);
  p.div_w1 = up_fastdiv((unsigned)(f->w + 1));
  p.div_h1 = up_fastdiv((unsigned)(f->h + 1));
  auto fit32 = [](long long v) { return (v >= -2147483647LL && v <= 2147483647LL) ? (int)v : kNeverLabel; };
  p.ignore32 = fit32(f->ignore_index);
  p.acc_ignore32 = f->acc_has_ignore ? fit32(f->acc_ignore_index) : kNeverLabel;
  if (p.cells == 0) return 0;
  B200SEG_REQUIRE(p.cells < (1LL << 31) && (long long)f->H * f->W < (1LL << 31) && (long long)f->N * f->C * f->h * f->w < (1LL << 31),
                  "loss_fused: problem too large for 32-bit cell / pixel / logit indices");
  // chunk width: 8 pixels when a run is at least ~6 pixels wide, else 4
  const bool wide = (long long)f->W >= 6LL * f->w;
  if (grad) return wide ? launch_upgen_px<T, 8, true>(p, st) : launch_upgen_px<T, 4, true>(p, st);
  return wide ? launch_upgen_px<T, 8, false>(p, st) : launch_upgen_px<T, 4, false>(p, st);
}

}  // namespace b200seg
