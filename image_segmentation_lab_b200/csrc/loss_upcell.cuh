// Resize-fused soft-max cross-entropy forward+backward, cell-owner formulation (sm_100a).
//
// Replaces, for logits at 1/S resolution (S a power of two in [4,32], align_corners=False, C <= 32), the chain
// resize (utils/ops.py:7-26) -> cross_entropy (models/losses/cross_entropy_loss.py:23-74) -> accuracy
// (models/losses/accuracy.py:6-61) and the autograd backward of all three, as called from
// models/decode_heads/decode_head.py:261-321, without materialising the (N,C,H,W) tensor in either direction.
//
// Geometry. Output pixel X has horizontal taps (r-1, r) with r = (X + S/2) / S and weight lx = (j + 0.5) / S for the
// j-th pixel of "run" r; rows likewise ("band" b, weight ly). The S x S pixels of cell (b, r) read exactly the 4
// low-resolution logits (b-1 | b) x (r-1 | r) per class and scatter their gradient to exactly those 4.
//
// Mapping. Four lanes ("quad") own one cell (x RG row groups for large S): lane q of the quad keeps classes
// [q*CPT, (q+1)*CPT) of the cell's 4 corner logits, the vertically interpolated tap pair of the current row, and the 4
// corner gradient sums of those classes IN REGISTERS for the whole cell. Pixels are walked in chunks of 4; every lane
// evaluates its classes for all 4 pixels (1 FFMA + 1 MUFU.EX2 + 1 FADD + 1/2 FMNMX per class-pixel forward, 2 FFMA
// backward, exponentials kept in registers between the two), partial sums cross the quad through shared memory once
// per chunk, and the per-pixel scalar work (label decode, log, reciprocal, loss, accuracy, one-hot term) is split so
// that lane q does it for pixel q of the chunk only. No CTA barrier, no atomics, no staging pass: the logits are
// read straight from L2/L1 (4 loads per class per cell), each cell's corner sums are written once to
// PB[n][c][band][run] (float4) and up_combine_kernel adds the 4 cells around every low-resolution logit.
// Deterministic (fixed summation order), unlike ATen's atomicAdd upsample backward.
//
// Numerics. Exponentials are taken against the row's upper bound M = max over classes of the two interpolated taps
// (an interpolated logit is a convex combination of its taps, so z - M <= 0): no per-pixel max pass. If a pixel's
// sum underflows (all its classes ~100 log2-units below M) the quad redoes that chunk against exact per-pixel maxima.
// Top-1: the label's class is the arg-max iff its exponential equals the maximum exponential of the pixel (ties among
// bit-identical interpolated logits count as correct for the label; torch.topk's choice among ties is unspecified).
//
// Bound: instruction issue / MUFU (C exponentials per output pixel); HBM traffic is the label map only.
// Algorithmic bytes per launch: 2*N*C*h*w*s + N*H*W*L.
#pragma once
#include "common.cuh"

namespace b200seg {

struct UpCellParams {
  const void* logits;
  const void* labels;
  const float* pw;
  const float* cw;
  unsigned long long* stats;
  float* pb;
  int label_dtype, label_bytes;
  int zero;           // always 0 (see the label prefetch)
  int has_w;          // class weights and/or per-pixel weights present (the rare path)
  int N, C, h, w, H, W;
  int S;
  int RG, logRG;      // lanes along the rows of one cell (row groups); each walks S / RG rows
  long long cells;    // N * (h + 1) * (w + 1)
  int ignore32;       // ignore_index / accuracy ignore_index as int32 (kNeverLabel if they do not fit: never matches)
  int acc_has_ignore;
  int acc_ignore32;
};

constexpr int kBigLabel = (int)0x80000000;     // a label value that does not fit int32 (never a class, never ignored)
constexpr int kNeverLabel = (int)0x80000001;   // an ignore value no decoded label can take

// Labels are consumed in the dtype the pipeline delivers. The raw word(s) are loaded one chunk ahead (the decode, which
// needs the data, happens at consumption) and squashed to int32: integer value if it fits, kBigLabel otherwise.
// LK: 0 = int64 (what the reference's label.long() delivers), 1 = uint8 (what datasets store), 2 = any (runtime dtype)
struct RawLabel { unsigned lo, hi; };
template <int LK> __device__ __forceinline__ RawLabel load_raw_label(const void* p, int dt, size_t i) {
  RawLabel r;
  r.hi = 0;
  if constexpr (LK == 0) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p) + i);
    r.lo = v.x; r.hi = v.y;
    return r;
  }
  if constexpr (LK == 1) {
    r.lo = __ldg(reinterpret_cast<const unsigned char*>(p) + i);
    return r;
  }
  if (dt == B200SEG_L_I64 || dt == B200SEG_L_F64) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p) + i);
    r.lo = v.x; r.hi = v.y;
  } else if (dt == B200SEG_L_U8) {
    r.lo = __ldg(reinterpret_cast<const unsigned char*>(p) + i);
  } else if (dt == B200SEG_L_I16) {
    r.lo = (unsigned)(int)__ldg(reinterpret_cast<const short*>(p) + i);
  } else {
    r.lo = __ldg(reinterpret_cast<const unsigned*>(p) + i);
  }
  return r;
}
template <int LK> __device__ __forceinline__ int decode_label(const RawLabel r, int dt) {
  if constexpr (LK == 0) return ((int)r.hi == ((int)r.lo >> 31)) ? (int)r.lo : kBigLabel;
  if constexpr (LK == 1) return (int)r.lo;
  if (dt == B200SEG_L_I64) return ((int)r.hi == ((int)r.lo >> 31)) ? (int)r.lo : kBigLabel;
  if (dt == B200SEG_L_F32 || dt == B200SEG_L_F64) {
    const double f = dt == B200SEG_L_F32 ? (double)__uint_as_float(r.lo) : __longlong_as_double(((long long)r.hi << 32) | r.lo);
    return (f > -2147483000.0 && f < 2147483000.0) ? (int)(long long)f : kBigLabel;
  }
  return (int)r.lo;
}

__device__ __forceinline__ float lg2(float x) {   // x is a normal number here (>= 1e-30)
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// class weight x per-pixel weight of one pixel: out of line so that the unweighted loop carries only a uniform branch
static __device__ __noinline__ float pixel_weight(const float* cw, const float* pw, bool use, int yc, size_t pix) {
  if (!use) return 0.f;
  float wt = cw ? __ldg(cw + yc) : 1.f;
  if (pw) wt *= __ldg(pw + pix);
  return wt;
}

constexpr int kXWords = 40;          // floats per quad in the exchange buffer (32 used; 40 keeps STS.64 conflict-free)
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kPadCorner = -1.0e30f;

template <int CPT, int THR> constexpr size_t upcell_smem_bytes(bool grad) {
  constexpr int kQuads = THR / 4, kCellThreads = THR;
  // LD [quad][4*CPT] float2 | X [quad][kXWords] float | INV [quad][4] float | OH [4*CPT][thread] float4 (GRAD)
  return (size_t)kQuads * 4 * CPT * 8 + (size_t)kQuads * kXWords * 4 + (size_t)kQuads * 16 +
         (grad ? (size_t)4 * CPT * kCellThreads * 16 : 0);
}

template <typename T, int CPT, bool GRAD, int THR, int MINB, int LK>
__global__ void __maxnreg__((65536 / (THR * MINB)) / 8 * 8 > 255 ? 255 : (65536 / (THR * MINB)) / 8 * 8) up_cell_kernel(const UpCellParams p) {
  constexpr int kCellThreads = THR, kQuads = THR / 4;
  constexpr int CT = 4 * CPT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* LD = reinterpret_cast<float2*>(smem_raw);                              // [kQuads][CT]
  float* XB = reinterpret_cast<float*>(LD + kQuads * CT);                        // [kQuads][kXWords]
  float* INV = XB + kQuads * kXWords;                                            // [kQuads][4]
  float4* OH = reinterpret_cast<float4*>(INV + kQuads * 4);                      // [CT][kCellThreads]

  const int tid = threadIdx.x, lane = tid & 31;
  const int q = lane & 3;
  const int quad = tid >> 2;
  const int rg = (lane >> 2) & (p.RG - 1);
  const int S = p.S, C = p.C;
  const int cells_per_warp = 8 >> p.logRG;
  const unsigned cid_raw = (blockIdx.x * (kCellThreads / 32) + (tid >> 5)) * cells_per_warp + (lane >> (2 + p.logRG));
  const bool cell_ok = cid_raw < (unsigned)p.cells;     // host: cells < 2^31
  const unsigned cid = cell_ok ? cid_raw : (unsigned)p.cells - 1u;
  const unsigned t0 = cid / (unsigned)(p.w + 1);
  const int r = (int)(cid - t0 * (unsigned)(p.w + 1));
  const int n = (int)(t0 / (unsigned)(p.h + 1));
  const int b = (int)(t0 - (unsigned)n * (unsigned)(p.h + 1));

  // ---- the cell's 4 corner logits for this lane's classes (classes >= C padded very negative: exp -> 0)
  float v00[CPT], dv0[CPT], v01[CPT], dv1[CPT];
  {
    const int plane = p.h * p.w;
    const T* pl = reinterpret_cast<const T*>(p.logits) + ((size_t)n * C + (size_t)(q * CPT)) * (size_t)plane;
    const int ya = b - 1 < 0 ? 0 : b - 1, yb = b > p.h - 1 ? p.h - 1 : b;
    const int xa = r - 1 < 0 ? 0 : r - 1, xb = r > p.w - 1 ? p.w - 1 : r;
    const int o00 = ya * p.w + xa, o01 = ya * p.w + xb, o10 = yb * p.w + xa, o11 = yb * p.w + xb;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      float a = kPadCorner, bq = kPadCorner, cq = kPadCorner, d = kPadCorner;
      if (q * CPT + c < C) {
        a = to_float<T>(__ldg(pl + o00));
        bq = to_float<T>(__ldg(pl + o01));
        cq = to_float<T>(__ldg(pl + o10));
        d = to_float<T>(__ldg(pl + o11));
      }
      pl += plane;
      v00[c] = a; v01[c] = bq; dv0[c] = cq - a; dv1[c] = d - bq;
    }
  }
  // Per-lane private corner sums OH[class][4 corners] in shared memory: this lane's classes receive the soft-max part at
  // the end of every row, any class receives the (negative) one-hot part of the pixels this lane does the scalars for.
  if constexpr (GRAD) {
#pragma unroll
    for (int c = 0; c < CT; ++c) OH[c * kCellThreads + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  // weights: lx = (j + 0.5) / S inside, constant 1 / 0 for the clamped first / last run (likewise ly for bands)
  const float invS = 1.f / (float)S;
  const float sx = (r == 0 || r == p.w) ? 0.f : invS, bx = (r == 0) ? 1.f : ((r == p.w) ? 0.f : 0.5f * invS);
  const float sy = (b == 0 || b == p.h) ? 0.f : invS, by = (b == 0) ? 1.f : ((b == p.h) ? 0.f : 0.5f * invS);
  float lxc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) lxc[j] = fmaf((float)j, sx, bx);
  const float lxcq = fmaf((float)q, sx, bx);
  const int rows_per = S >> p.logRG;
  const int Xbase = S * r - S / 2;
  float loss_acc = 0.f;
  int n_valid = 0, n_correct = 0, n_bad = 0, n_acc = 0;
  float2* ldq = LD + quad * CT;
  float* xq = XB + quad * kXWords;
  const unsigned gmask = 0xFu << (lane & 28);

  // Label addressing: per-image byte pointer + 32-bit in-image offsets (host: H * W < 2^31); coordinates are clamped so
  // every load is unconditional, validity only masks the results.
  const int Y0 = S * b - S / 2 + rg * rows_per;
  const int dt = p.label_dtype;
  const int lb = LK == 0 ? 8 : (LK == 1 ? 1 : p.label_bytes);
  const size_t img_px = (size_t)n * p.H * p.W;
  const char* labimg = reinterpret_cast<const char*>(p.labels) + img_px * lb;
  const int Xq = Xbase + q;
  auto clampx = [&](int X) { return min(max(X, 0), p.W - 1); };
  auto clampy = [&](int Y) { return min(max(Y, 0), p.H - 1); };
  RawLabel raw_next = load_raw_label<LK>(labimg, dt, (unsigned)(clampy(Y0) * p.W + clampx(Xq)));

#pragma unroll 1
  for (int ii = 0; ii < rows_per; ++ii) {
    const int i = rg * rows_per + ii;
    const int Y = S * b - S / 2 + i;
    const bool row_ok = cell_ok && Y >= 0 && Y < p.H;
    const float ly = fmaf((float)i, sy, by);
    // ---- row prologue: vertical interpolation of the tap pair, the row's upper bound, scaled taps for ex2
    float L2[CPT], D2[CPT];
    {
      float Lr[CPT], Rr[CPT];
      float mloc = kPadCorner;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        Lr[c] = fmaf(ly, dv0[c], v00[c]);
        Rr[c] = fmaf(ly, dv1[c], v01[c]);
        mloc = fmaxf(mloc, fmaxf(Lr[c], Rr[c]));
      }
      mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, 1));
      mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, 2));
      const float nM2 = -mloc * kLog2e;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        L2[c] = fmaf(Lr[c], kLog2e, nM2);
        D2[c] = (Rr[c] - Lr[c]) * kLog2e;
        ldq[q * CPT + c] = make_float2(L2[c], D2[c]);
      }
    }
    float gs[CPT], gb[CPT];
    if constexpr (GRAD) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) { gs[c] = 0.f; gb[c] = 0.f; }
    }
    __syncwarp();
    const int roff = clampy(Y) * p.W, roff1 = clampy(Y + 1) * p.W;
    {   // pull the labels of two rows ahead into L2 (one sector per chunk of an int64 map)
      const int roff2 = clampy(Y + 2) * p.W;
      for (int jc = 4 * q; jc < S; jc += 16)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(labimg + (size_t)(unsigned)(roff2 + clampx(Xbase + jc)) * lb));
    }

#pragma unroll 1
    for (int j0 = 0; j0 < S; j0 += 4) {
      // ---- own pixel (pixel q of the chunk): label (loaded one chunk ahead), weight, counters
      const int Xo = Xq + j0;
      const int Xc = clampx(Xo);
      const bool pok = row_ok && Xo == Xc;
      // The label was loaded one chunk ahead. It is decoded BEFORE the next load is issued, and the next load's address
      // depends on it (`ydec & p.zero` == 0, opaque to the compiler): the two loads share a scoreboard slot, so a wait
      // placed after the new load would wait for the new load (measured: 18 % of all stall samples).
      const int ydec = decode_label<LK>(raw_next, dt);
      {
        const bool last = (j0 + 4 == S);
        const int Xn = clampx(Xq + (last ? 0 : j0 + 4));
        raw_next = load_raw_label<LK>(labimg, dt, (unsigned)((last ? roff1 : roff) + Xn) + (unsigned)(ydec & p.zero));
      }
      const int yy = pok ? ydec : p.ignore32;
      const bool ign = (yy == p.ignore32);
      const bool inr = (unsigned)yy < (unsigned)C;
      const int yc = inr && !ign ? yy : 0;
      const bool use = pok && inr && !ign;
      const bool acc_ok = pok && (p.acc_has_ignore ? (yy != p.acc_ignore32) : true);
      float wt = use ? 1.f : 0.f;
      if (p.has_w) wt = pixel_weight(p.cw, p.pw, use, yc, img_px + (size_t)(unsigned)(roff + Xc));
      n_valid += (pok && !ign);
      n_bad += (pok && !ign && !inr);
      n_acc += acc_ok;
      // ---- horizontal weights of the chunk's 4 pixels
      const float fj0 = (float)j0;
      float lx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) lx[j] = fmaf(fj0, sx, lxc[j]);
      const float lxo = fmaf(fj0, sx, lxcq);
      // ---- the label's interpolated logit (any lane's class: read from the quad's row buffer), before the class loop so
      // that its shared-memory latency is off the exchange -> loss chain
      const float2 ldy = ldq[yc];
      const float zy2 = fmaf(lxo, ldy.y, ldy.x);   // same operation as the class loop: bitwise the label's z
      float ey = ex2(zy2);
      // ---- class loop, forward: exponentials against the row bound, partial sums and maxima of this lane's classes
      float e[4][CPT];
      float sp[4], mp[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { sp[j] = 0.f; mp[j] = 0.f; }
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float ev = ex2(fmaf(lx[j], D2[c], L2[c]));
          e[j][c] = ev;
          sp[j] += ev;
          mp[j] = fmaxf(mp[j], ev);
        }
      }
      // ---- quad exchange: lane q receives the 4 partials of pixel q
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float2*>(xq + j * 8 + q * 2) = make_float2(sp[j], mp[j]);
      __syncwarp();
      float s, em;
      {
        const float4 u0 = *reinterpret_cast<const float4*>(xq + q * 8);
        const float4 u1 = *reinterpret_cast<const float4*>(xq + q * 8 + 4);
        s = (u0.x + u0.z) + (u1.x + u1.z);
        em = fmaxf(fmaxf(u0.y, u0.w), fmaxf(u1.y, u1.w));
      }
      float mofs = 0.f;
      // ---- rare: a sum underflowed against the row bound -> redo the quad's chunk against exact per-pixel maxima
      const unsigned under = __ballot_sync(0xffffffffu, !(s > 1e-30f));
      if (under) {
        if ((under >> (lane & 28)) & 0xFu) {
          float mz[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float m = -3.0e38f;
#pragma unroll
            for (int c = 0; c < CPT; ++c) m = fmaxf(m, fmaf(lx[j], D2[c], L2[c]));
            m = fmaxf(m, __shfl_xor_sync(gmask, m, 1));
            m = fmaxf(m, __shfl_xor_sync(gmask, m, 2));
            mz[j] = m;
            float sj = 0.f, mj = 0.f;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              const float ev = ex2(fmaf(lx[j], D2[c], L2[c]) - m);
              e[j][c] = ev;
              sj += ev;
              mj = fmaxf(mj, ev);
            }
            sj += __shfl_xor_sync(gmask, sj, 1);
            sj += __shfl_xor_sync(gmask, sj, 2);
            mj = fmaxf(mj, __shfl_xor_sync(gmask, mj, 1));
            mj = fmaxf(mj, __shfl_xor_sync(gmask, mj, 2));
            sp[j] = sj;
            mp[j] = mj;
          }
          s = q == 0 ? sp[0] : (q == 1 ? sp[1] : (q == 2 ? sp[2] : sp[3]));
          em = q == 0 ? mp[0] : (q == 1 ? mp[1] : (q == 2 ? mp[2] : mp[3]));
          mofs = q == 0 ? mz[0] : (q == 1 ? mz[1] : (q == 2 ? mz[2] : mz[3]));
          ey = ex2(zy2 - mofs);
        }
      }
      // ---- per-pixel scalars of the own pixel (log2 units; ln 2 is applied once per thread)
      loss_acc = fmaf(wt, (mofs + lg2(s)) - zy2, loss_acc);
      n_correct += (acc_ok && inr && !ign && ey == em);
      if constexpr (GRAD) {
        INV[quad * 4 + q] = wt * fast_rcp(s);
        __syncwarp();
        const float4 iv = *reinterpret_cast<const float4*>(INV + quad * 4);
        const float ia[4] = {iv.x, iv.y, iv.z, iv.w};
        // ---- class loop, backward: horizontal corner sums of wt * softmax
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = ia[j], bb = ia[j] * lx[j];
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            gs[c] = fmaf(e[j][c], a, gs[c]);
            gb[c] = fmaf(e[j][c], bb, gb[c]);
          }
        }
        // ---- one-hot term of the own pixel into this lane's private per-class corner sums
        {
          const float u = wt * lxo, v = wt - u;
          float4* oh = OH + yc * kCellThreads + tid;
          float4 o = *oh;
          o.x = fmaf(ly - 1.f, v, o.x);
          o.y = fmaf(ly - 1.f, u, o.y);
          o.z = fmaf(-ly, v, o.z);
          o.w = fmaf(-ly, u, o.w);
          *oh = o;
        }
      } else {
        __syncwarp();   // XB is rewritten by the next chunk
      }
    }
    if constexpr (GRAD) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const float ga = gs[c] - gb[c];
        float4* oh = OH + (q * CPT + c) * kCellThreads + tid;
        float4 o = *oh;
        o.x = fmaf(1.f - ly, ga, o.x);
        o.y = fmaf(1.f - ly, gb[c], o.y);
        o.z = fmaf(ly, ga, o.z);
        o.w = fmaf(ly, gb[c], o.w);
        *oh = o;
      }
    }
    __syncwarp();   // LD is rewritten by the next row
  }

  if constexpr (GRAD) {
    __syncwarp();
    const int qbase = tid & ~3;
    float acc[CPT][4];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const float4* oh = OH + (q * CPT + c) * kCellThreads + qbase;
      const float4 o0 = oh[0], o1 = oh[1], o2 = oh[2], o3 = oh[3];
      acc[c][0] = (o0.x + o1.x) + (o2.x + o3.x);
      acc[c][1] = (o0.y + o1.y) + (o2.y + o3.y);
      acc[c][2] = (o0.z + o1.z) + (o2.z + o3.z);
      acc[c][3] = (o0.w + o1.w) + (o2.w + o3.w);
    }
    for (int off = 4; off < (4 << p.logRG); off <<= 1) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[c][k] += __shfl_xor_sync(0xffffffffu, acc[c][k], off);
      }
    }
    if (cell_ok && rg == 0) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const int cg = q * CPT + c;
        if (cg < C) {
          float4* dst = reinterpret_cast<float4*>(p.pb) + (((size_t)n * C + cg) * (p.h + 1) + b) * (p.w + 1) + r;
          *dst = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
        }
      }
    }
  }
  cta_flush_stats(loss_acc * kLn2, n_valid, n_correct, n_bad, n_acc, p.stats);
}

template <typename T, int CPT, bool GRAD, int THR, int MINB, int LK> static int launch_upcell_lk(const UpCellParams& p, cudaStream_t st) {
  constexpr size_t smem = upcell_smem_bytes<CPT, THR>(GRAD);
  constexpr int kCellThreads = THR;
  auto k = up_cell_kernel<T, CPT, GRAD, THR, MINB, LK>;
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(k), (int)smem)) return e;
  const long long cells_per_cta = (long long)(kCellThreads / 32) * (8 >> p.logRG);
  const long long grid = (p.cells + cells_per_cta - 1) / cells_per_cta;   // one cell group per warp: measured faster
                                                                          // than persistent warps with next-cell prefetch
  k<<<(unsigned)grid, kCellThreads, smem, st>>>(p);
  count_launch();
  return check_launch("up_cell_kernel");
}

template <typename T, int CPT, bool GRAD, int THR, int MINB> static int launch_upcell_mb(const UpCellParams& p, cudaStream_t st) {
  if (p.label_dtype == B200SEG_L_I64) return launch_upcell_lk<T, CPT, GRAD, THR, MINB, 0>(p, st);
  if (p.label_dtype == B200SEG_L_U8) return launch_upcell_lk<T, CPT, GRAD, THR, MINB, 1>(p, st);
  return launch_upcell_lk<T, CPT, GRAD, THR, MINB, 2>(p, st);
}

// CTAs of 64 threads (16 cells): small CTAs keep the tail of the last wave short. Register cap per thread from the
// resident CTAs per SM: 8 x 64 threads -> 128, 6 -> 168, 4 -> 255. Measured on B200 for C = 19 (CPT 5): 6 CTAs/SM
// without spills (80 us) beat 7 and 8 CTAs/SM with a tighter cap (85 / 87 us).
template <typename T, int CPT, bool GRAD> static int launch_upcell(const UpCellParams& p, cudaStream_t st) {
  constexpr int MINB = CPT <= 3 ? 8 : (CPT <= 6 ? 6 : 4);
  return launch_upcell_mb<T, CPT, GRAD, 64, MINB>(p, st);
}

template <typename T, bool GRAD> static int pick_upcell(const UpCellParams& p, cudaStream_t st) {
  switch ((p.C + 3) / 4) {
    case 1: return launch_upcell<T, 1, GRAD>(p, st);
    case 2: return launch_upcell<T, 2, GRAD>(p, st);
    case 3: return launch_upcell<T, 3, GRAD>(p, st);
    case 4: return launch_upcell<T, 4, GRAD>(p, st);
    case 5: return launch_upcell<T, 5, GRAD>(p, st);
    case 6: return launch_upcell<T, 6, GRAD>(p, st);
    case 7: return launch_upcell<T, 7, GRAD>(p, st);
    default: return launch_upcell<T, 8, GRAD>(p, st);
  }
}

// Row groups per cell: enough lanes to fill the machine when the cells are few and large (S = 16, 32).
static inline int pick_row_groups(long long cells, int S) {
  int rgv = 1;
  const long long want = (long long)kSMs * 4 * 128 * 3;   // >= 3 rounds of resident CTAs
  while (rgv < 8 && rgv * 2 <= S / 2 && cells * 4 * rgv < want) rgv *= 2;
  return rgv;
}

template <typename T> int upcell_run(const b200seg_loss_desc* f, float* pb, int S, bool grad, cudaStream_t st) {
  UpCellParams p;
  p.logits = f->logits; p.labels = f->labels; p.pw = f->pixel_weight; p.cw = f->ce_class_weight;
  p.stats = reinterpret_cast<unsigned long long*>(f->stats);
  p.pb = pb;
  p.label_dtype = f->label_dtype; p.label_bytes = label_bytes(f->label_dtype);
  p.has_w = (p.cw != nullptr) || (p.pw != nullptr);
  p.zero = 0;
  p.N = f->N; p.C = f->C; p.h = f->h; p.w = f->w; p.H = f->H; p.W = f->W;
  p.S = S;
  p.cells = (long long)f->N * (f->h + 1) * (f->w + 1);
  p.RG = pick_row_groups(p.cells, S);
  p.logRG = 0; while ((1 << p.logRG) < p.RG) ++p.logRG;
  auto fit32 = [](long long v) { return (v >= -2147483647LL && v <= 2147483647LL) ? (int)v : kNeverLabel; };
  p.ignore32 = fit32(f->ignore_index); p.acc_has_ignore = f->acc_has_ignore; p.acc_ignore32 = fit32(f->acc_ignore_index);
  if (p.cells == 0) return 0;
  B200SEG_REQUIRE(p.cells < (1LL << 31) && (long long)f->H * f->W < (1LL << 31), "loss_fused: problem too large for 32-bit cell / pixel indices");
  return grad ? pick_upcell<T, true>(p, st) : pick_upcell<T, false>(p, st);
}


}  // namespace b200seg
