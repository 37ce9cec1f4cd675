// Single-pass forward+backward of resize + softmax cross-entropy (+ accuracy) on LOW-RESOLUTION logits, sm_100a.
//
// For a 'mean' / 'sum' reduction the gradient of the loss w.r.t. a pixel's logits depends on that pixel only (up to
// one global scale), so it is produced while the interpolated logits are still in registers: the (N,C,H,W) tensor is
// never materialised in either direction, instead of the reference's ~10 passes over it (utils/ops.py:26,
// cross_entropy_loss.py:56-61, accuracy.py:41 and their autograd backwards).
//
// Scope: logits at 1/S resolution, S a power of two in [4,32], align_corners=False (the decode-head call,
// models/decode_heads/decode_head.py:266-269), C <= 32. Other ratios take the general kernels of loss_stream.cu.
//
// Geometry: output pixel X has taps (r-1, r) with r = (X + S/2) / S; the S consecutive pixels of one "run" r share
// their taps, likewise the S rows of a "band" b. A (band, run) cell of SxS output pixels therefore scatters its
// gradient to exactly 4 low-res corners. A CTA (256 threads) owns one band x a range of runs: thread = (row of the
// band, 4 consecutive pixels of one run). It builds the C interpolated logits of its 4 pixels in registers (1 FFMA
// each from the vertically interpolated tap pair read from a shared-memory patch with compile-time strides), does
// the soft-max (one MUFU.EX2 per element) and reduces its gradient to the two horizontal corners. The per-cell
// reduction over S rows x S/4 threads goes through shared memory in a FIXED order; each cell's 4 corner sums are
// written once to PB[n][c][band][run][2][2]; up_combine_kernel adds the 4 cells around every low-res logit and
// applies the global scale (upstream gradient, loss_weight, 1/denominator). No atomics: the backward is
// deterministic (ATen's upsample_bilinear2d_backward is an atomicAdd scatter).
//
// Bound: instruction issue (C ex2 + ~14 C FP32/ALU instructions per output pixel); the only full-resolution tensor
// touched is the label map. Algorithmic bytes per launch: 2*N*C*h*w*s + N*H*W*L.
#include "common.cuh"

namespace b200seg {

constexpr int kPatchStride = 68;   // floats per patch row (>= runs per tile + 1 = 65 at S = 4)
constexpr float kPadLogit = -1.0e30f;
#ifndef B200SEG_UP_MINBLOCKS
#define B200SEG_UP_MINBLOCKS 4
#endif
constexpr int kUpMinBlocks = B200SEG_UP_MINBLOCKS;   // resident CTAs per SM the S >= 8 variant is compiled for

struct UpParams {
  const void* logits;
  const void* labels;
  const float* pw;
  const float* cw;
  unsigned long long* stats;
  float* pb;
  int label_dtype;
  int N, C, h, w, H, W;
  int S;
  int NG;    // thread groups (4 px) per row per CTA = 256 / S
  int GPR;   // groups per run = S / 4
  int RT;    // runs per tile = NG / GPR
  int logNG, logGPR, logRT;
  long long ignore_index;
  int acc_has_ignore;
  long long acc_ignore;
};

// Shared-memory layout of one CTA (floats):
//   raw   [CPT][2][kPatchStride]   the two clamped low-res tap rows of the band (classes >= C padded very negative)
//   vpat  [CPT][kVStride]          vertically interpolated taps V[c][row i][col k] = a + ly_i (b - a), index i*ncol + k
//   mrow  [kVStride]               max over classes of V[.][i][k]: an upper bound of every interpolated logit
//   stage [CPT][256] float2        (GRAD) per-thread horizontal corner sums; aliases raw
// VS = floats per class of vpat: >= S * (RT + 1), i.e. 260 at S = 4 and 136 / 80 / 64 at S = 8 / 16 / 32

template <typename T, int CPT, bool GRAD, int VS, int MINB>
__global__ void __launch_bounds__(256, MINB) up_fused_kernel(const UpParams p) {
  constexpr int kVStride = VS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float lam_y[32];  // cell-relative vertical weight of each row of the band (-1: row outside the image)
  const int C = p.C, S = p.S, NG = p.NG, GPR = p.GPR, RT = p.RT;
  const int n = blockIdx.z, b = blockIdx.y, tile = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r_first = tile * RT;
  const int ncol = RT + 1;

  // raw is dead once vpat is built, so it shares its space with stage (first used after the next barrier)
  constexpr int kUnionFloats = (GRAD && CPT * 512 > CPT * 2 * kPatchStride) ? CPT * 512 : CPT * 2 * kPatchStride;
  float* raw = reinterpret_cast<float*>(smem_raw);
  float2* stage = reinterpret_cast<float2*>(smem_raw);
  float* vpat = raw + kUnionFloats;
  float* mrow = vpat + CPT * kVStride;

  // ---- stage the two low-res tap rows of this band (clamped) as fp32; classes >= C are padded very negative
  {
    const T* img = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * p.h * p.w;
    const int ya = b - 1 < 0 ? 0 : b - 1, yb = b > p.h - 1 ? p.h - 1 : b;
    for (int c = warp; c < CPT; c += 8) {
      for (int t = lane; t < 2 * ncol; t += 32) {
        const int rr = t >= ncol ? 1 : 0;
        const int k = t - rr * ncol;
        float v = kPadLogit;
        if (c < C) {
          int col = r_first - 1 + k;
          col = col < 0 ? 0 : (col > p.w - 1 ? p.w - 1 : col);
          v = to_float<T>(img[((size_t)c * p.h + (rr ? yb : ya)) * p.w + col]);
        }
        raw[(c * 2 + rr) * kPatchStride + k] = v;
      }
    }
    if (tid < S) {
      const int Y = S * b - S / 2 + tid;
      float l;
      if (b == 0) l = 1.f;
      else if (b == p.h) l = 0.f;
      else l = ((float)Y + 0.5f) / (float)S - 0.5f - (float)(b - 1);  // exact for power-of-two S
      lam_y[tid] = (Y >= 0 && Y < p.H) ? l : -1.f;
    }
  }
  __syncthreads();
  // ---- vertical interpolation once per (row, column, class) for the whole CTA, and the per-(row, column) class max
  for (int e = tid; e < S * ncol; e += 256) {
    const int i = e / ncol, k = e - i * ncol;
    const float l = lam_y[i] < 0.f ? 0.f : lam_y[i];
    float mx = kPadLogit;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const float a = raw[(c * 2 + 0) * kPatchStride + k], bb = raw[(c * 2 + 1) * kPatchStride + k];
      const float v = fmaf(l, bb - a, a);
      vpat[c * kVStride + e] = v;
      mx = fmaxf(mx, v);
    }
    mrow[e] = mx;
  }
  __syncthreads();

  const int i = tid >> p.logNG;  // row within the band
  const int ul = tid & (NG - 1); // group within the tile
  const int Y = S * b - S / 2 + i;
  const int u = tile * NG + ul;
  const int X0 = 4 * u - S / 2;
  const int r = u >> p.logGPR;
  const bool row_ok = (Y >= 0 && Y < p.H);
  const bool any_ok = row_ok && r <= p.w && X0 + 3 >= 0 && X0 < p.W;

  float loss_acc = 0.f;
  int n_valid = 0, n_correct = 0, n_bad = 0, n_acc = 0;

  if (any_ok) {
    // horizontal weights of the 4 pixels: lx[j] = lx0 + j/S (exact for power-of-two S); runs 0 / w are clamped
    float lx[4];
    bool pix_ok[4];
    {
      const float invS = 1.f / (float)S;
      const float lx0 = ((float)X0 + 0.5f) * invS - 0.5f - (float)(r - 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int X = X0 + j;
        pix_ok[j] = (X >= 0 && X < p.W);
        lx[j] = (r == 0) ? 1.f : ((r == p.w) ? 0.f : fmaf((float)j, invS, lx0));
      }
    }
    // labels of the 4 pixels as 32-bit class indices: >= 0 valid, -1 not in [0,C) ("bad"), -2 ignored
    int y32[4];
    bool acc_ok[4];
    const size_t lbase = ((size_t)n * p.H + Y) * p.W;
    if (p.label_dtype == B200SEG_L_I64 && X0 >= 0 && X0 + 3 < p.W && ((lbase + X0) & 1) == 0 && aligned16(p.labels)) {
      const char* lp = reinterpret_cast<const char*>(p.labels) + (lbase + X0) * 8;
      const uint4 a = ld_stream16(lp), c = ld_stream16(lp + 16);
      const unsigned lo[4] = {a.x, a.z, c.x, c.z}, hi[4] = {a.y, a.w, c.y, c.w};
      const unsigned ig_lo = (unsigned)((unsigned long long)p.ignore_index), ig_hi = (unsigned)((unsigned long long)p.ignore_index >> 32);
      const unsigned ag_lo = (unsigned)((unsigned long long)p.acc_ignore), ag_hi = (unsigned)((unsigned long long)p.acc_ignore >> 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ign = (lo[j] == ig_lo) & (hi[j] == ig_hi);
        const bool inr = (hi[j] == 0u) & (lo[j] < (unsigned)C);
        y32[j] = ign ? -2 : (inr ? (int)lo[j] : -1);
        acc_ok[j] = p.acc_has_ignore ? !((lo[j] == ag_lo) & (hi[j] == ag_hi)) : true;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long yy = pix_ok[j] ? load_label(p.labels, p.label_dtype, lbase + X0 + j) : p.ignore_index;
        y32[j] = (yy == p.ignore_index) ? -2 : ((yy >= 0 && yy < (long long)C) ? (int)yy : -1);
        acc_ok[j] = p.acc_has_ignore ? (yy != p.acc_ignore) : true;
      }
    }

    // ---- pass 1: interpolated logits, arg-max, and sum of exponentials against the reference M >= every logit of
    // these 4 pixels (class max of the two tap columns: a logit is a convex combination of its taps)
    const float* pv = vpat + i * ncol + (r - r_first);
    const float M = fmaxf(mrow[i * ncol + (r - r_first)], mrow[i * ncol + (r - r_first) + 1]);
    const float nM = -M * kLog2e;
    float m[4], s[4];
    int idx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { m[j] = neg_inf(); idx[j] = 0; s[j] = 0.f; }
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const float va = pv[c * kVStride], vb = pv[c * kVStride + 1];
      const float d = vb - va;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float z = fmaf(lx[j], d, va);
        if (z > m[j]) { m[j] = z; idx[j] = c; }   // strict '>' keeps the lowest index
        s[j] += ex2(fmaf(z, kLog2e, nM));
      }
    }
    // re-reference the sums to each pixel's own max (s >= 1 afterwards). If a pixel sits more than ~80 below the tile
    // bound its sum underflowed: redo that (rare) pixel exactly against its own max.
    bool redo = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) redo |= !(s[j] > 1e-30f);
    if (redo) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] = 0.f;
      for (int c = 0; c < C; ++c) {
        const float va = pv[c * kVStride], vb = pv[c * kVStride + 1];
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] += ex2((fmaf(lx[j], vb - va, va) - m[j]) * kLog2e);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] *= ex2((M - m[j]) * kLog2e);
    }

    float coef[4];
    int ycl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      coef[j] = 0.f;
      ycl[j] = -1;
      if (pix_ok[j]) {
        const int yc = y32[j];
        n_bad += (yc == -1);
        n_valid += (yc != -2);
        if (yc >= 0) {
          const float va = pv[yc * kVStride], vb = pv[yc * kVStride + 1];
          const float zy = fmaf(lx[j], vb - va, va);   // same operations as the class loop: bitwise equal
          float wt = p.cw ? __ldg(p.cw + yc) : 1.f;
          if (p.pw) wt *= __ldg(p.pw + lbase + X0 + j);
          loss_acc = fmaf(wt, m[j] + fast_log(s[j]) - zy, loss_acc);
          coef[j] = wt;
          ycl[j] = yc;
        }
        n_acc += acc_ok[j];
        n_correct += (acc_ok[j] && idx[j] == yc);
      }
    }
    if constexpr (GRAD) {
      // ---- pass 2: the exponentials are recomputed (not kept: 80 registers), now against each pixel's own max
      float rj[4], nm[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { rj[j] = coef[j] * fast_rcp(s[j]); nm[j] = -m[j] * kLog2e; }
      float2* st = stage + tid;   // i * NG + ul == tid
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const float va = pv[c * kVStride], vb = pv[c * kVStride + 1];
        const float d = vb - va;
        float gs = 0.f, gb = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float g = rj[j] * ex2(fmaf(fmaf(lx[j], d, va), kLog2e, nm[j]));
          gs += g;
          gb = fmaf(lx[j], g, gb);
        }
        st[c * 256] = make_float2(gs - gb, gb);
      }
      // one-hot part: subtract coef at the label class (own slot: plain read-modify-write)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ycl[j] >= 0) {
          float2 v = st[ycl[j] * 256];
          v.x -= (1.f - lx[j]) * coef[j];
          v.y -= lx[j] * coef[j];
          st[ycl[j] * 256] = v;
        }
      }
    }
  } else if constexpr (GRAD) {
    float2* st = stage + tid;
#pragma unroll
    for (int c = 0; c < CPT; ++c) st[c * 256] = make_float2(0.f, 0.f);
  }

  if constexpr (GRAD) {
    __syncthreads();
    // fixed-order per-cell reduction: thread = run of the tile (both corner rows), classes strided over the CTA
    const int rl = tid & (RT - 1);
    const int rr = r_first + rl;
    if (rr <= p.w) {
      const int cstep = 256 >> p.logRT;
      for (int c = tid >> p.logRT; c < C; c += cstep) {
        float s0a = 0.f, s0b = 0.f, s1a = 0.f, s1b = 0.f;
        const float2* base = stage + c * 256 + rl * GPR;
        for (int ii = 0; ii < S; ++ii) {
          const float l = lam_y[ii];
          if (l < 0.f) continue;
          const float2* row = base + ii * NG;
          float ra = 0.f, rb = 0.f;
          for (int q = 0; q < GPR; ++q) { ra += row[q].x; rb += row[q].y; }
          s1a = fmaf(l, ra, s1a);
          s1b = fmaf(l, rb, s1b);
          s0a = fmaf(1.f - l, ra, s0a);
          s0b = fmaf(1.f - l, rb, s0b);
        }
        float4* dst = reinterpret_cast<float4*>(p.pb) + (((size_t)n * C + c) * (p.h + 1) + b) * (p.w + 1) + rr;
        *dst = make_float4(s0a, s0b, s1a, s1b);
      }
    }
  }
  cta_flush_stats(loss_acc, n_valid, n_correct, n_bad, n_acc, p.stats);
}

// grad[n][c][y][x] = G * (cell(y,x).c11 + cell(y,x+1).c10 + cell(y+1,x).c01 + cell(y+1,x+1).c00)
template <typename T>
__global__ void __launch_bounds__(256) up_combine_kernel(const float* __restrict__ pb, T* __restrict__ grad, int NC, int h,
                                                         int w, float scale_host, const float* grad_out, int use_nvalid,
                                                         const unsigned long long* stats) {
  float G = scale_host;
  if (grad_out) G *= __ldg(grad_out);
  if (use_nvalid) {
    const double nv = (double)(long long)stats[B200SEG_ST_N_VALID];
    G = (float)((double)G / (nv + 1.1920928955078125e-07));
  }
  const long long total = (long long)NC * h * w;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % w);
    const long long t = idx / w;
    const int y = (int)(t % h);
    const long long nc = t / h;
    const float4* cells = reinterpret_cast<const float4*>(pb) + nc * (long long)(h + 1) * (w + 1);
    const float4 c00 = cells[(long long)y * (w + 1) + x];            // cell (b=y,   r=x)   -> corner (1,1) = .w
    const float4 c01 = cells[(long long)y * (w + 1) + x + 1];        // cell (b=y,   r=x+1) -> corner (1,0) = .z
    const float4 c10 = cells[(long long)(y + 1) * (w + 1) + x];      // cell (b=y+1, r=x)   -> corner (0,1) = .y
    const float4 c11 = cells[(long long)(y + 1) * (w + 1) + x + 1];  // cell (b=y+1, r=x+1) -> corner (0,0) = .x
    grad[idx] = from_float<T>(G * ((c00.w + c01.z) + (c10.y + c11.x)));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scale_inplace_kernel(T* __restrict__ x, long long n, const float* __restrict__ g) {
  const float s = __ldg(g);
  if (s == 1.f) return;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = from_float<T>(to_float<T>(x[i]) * s);
}

static bool up_fast_ok(int C, int h, int w, int H, int W, int ac, int* S_out) {
  if (ac) return false;
  if (h <= 0 || w <= 0 || H % h || W % w) return false;
  const int S = H / h;
  if (S != W / w || S < 4 || S > 32 || (S & (S - 1))) return false;
  if (C > 32) return false;
  *S_out = S;
  return true;
}

long long up_fused_workspace(int N, int C, int h, int w, int H, int W, int ac) {
  int S;
  if (!up_fast_ok(C, h, w, H, W, ac, &S)) return 0;
  return (long long)N * C * (h + 1) * (w + 1) * 4 * (long long)sizeof(float);
}

template <typename T, int CPT, bool GRAD, int VS> static int launch_up_vs(const UpParams& p, cudaStream_t st) {
  constexpr int kVStride = VS;
  constexpr int MINB = (VS <= 136 && CPT <= 24) ? kUpMinBlocks : 3;
  const size_t uni = (GRAD && CPT * 512 > CPT * 2 * kPatchStride) ? (size_t)CPT * 512 : (size_t)CPT * 2 * kPatchStride;
  const size_t smem = (uni + (size_t)CPT * kVStride + kVStride) * 4;
  auto k = up_fused_kernel<T, CPT, GRAD, VS, MINB>;
  static bool attr = false;
  if (!attr) {
    B200SEG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int NU = (p.W + p.S / 2 + 3) / 4;
  dim3 grid((NU + p.NG - 1) / p.NG, p.h + 1, p.N);
  k<<<grid, 256, smem, st>>>(p);
  count_launch();
  return check_launch("up_fused_kernel");
}

template <typename T, int CPT, bool GRAD> static int launch_up(const UpParams& p, cudaStream_t st) {
  return p.S == 4 ? launch_up_vs<T, CPT, GRAD, 264>(p, st) : launch_up_vs<T, CPT, GRAD, 136>(p, st);
}

template <typename T, bool GRAD> static int pick_up(const UpParams& p, cudaStream_t st) {
  if (p.C <= 4) return launch_up<T, 4, GRAD>(p, st);
  if (p.C <= 8) return launch_up<T, 8, GRAD>(p, st);
  if (p.C <= 12) return launch_up<T, 12, GRAD>(p, st);
  if (p.C <= 16) return launch_up<T, 16, GRAD>(p, st);
  if (p.C <= 20) return launch_up<T, 20, GRAD>(p, st);
  if (p.C <= 24) return launch_up<T, 24, GRAD>(p, st);
  return launch_up<T, 32, GRAD>(p, st);
}

template <typename T>
static int up_combine_t(const void* ws, void* grad, int N, int C, int h, int w, float scale_host, const float* grad_out,
                        int use_nvalid, const uint64_t* stats, cudaStream_t st) {
  const long long total = (long long)N * C * h * w;
  long long blocks = (total + 255) / 256;
  if (blocks > kSMs * 8) blocks = kSMs * 8;
  if (blocks < 1) blocks = 1;
  up_combine_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float*>(ws), reinterpret_cast<T*>(grad),
                                                        N * C, h, w, scale_host, grad_out, use_nvalid,
                                                        reinterpret_cast<const unsigned long long*>(stats));
  count_launch();
  return check_launch("up_combine_kernel");
}

int up_combine_dispatch(const void* ws, void* grad, int dtype, int N, int C, int h, int w, float scale_host,
                        const float* grad_out, int use_nvalid, const uint64_t* stats, cudaStream_t st) {
  switch (dtype) {
    case B200SEG_F32: return up_combine_t<float>(ws, grad, N, C, h, w, scale_host, grad_out, use_nvalid, stats, st);
    case B200SEG_BF16: return up_combine_t<__nv_bfloat16>(ws, grad, N, C, h, w, scale_host, grad_out, use_nvalid, stats, st);
    case B200SEG_F16: return up_combine_t<__half>(ws, grad, N, C, h, w, scale_host, grad_out, use_nvalid, stats, st);
  }
  set_error("loss_fused_combine: unsupported dtype %d", dtype);
  return 1;
}

int scale_inplace_dispatch(void* x, int dtype, long long n, const float* g, cudaStream_t st) {
  long long blocks = (n + 255) / 256;
  if (blocks > kSMs * 16) blocks = kSMs * 16;
  if (blocks < 1) blocks = 1;
  switch (dtype) {
    case B200SEG_F32: scale_inplace_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((float*)x, n, g); break;
    case B200SEG_BF16: scale_inplace_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((__nv_bfloat16*)x, n, g); break;
    case B200SEG_F16: scale_inplace_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>((__half*)x, n, g); break;
    default: set_error("scale_inplace: unsupported dtype %d", dtype); return 1;
  }
  count_launch();
  return check_launch("scale_inplace_kernel");
}

template <typename T> static int up_run(const b200seg_loss_fused_desc* d, int S, cudaStream_t st) {
  const b200seg_loss_desc* f = &d->fwd;
  UpParams p;
  p.logits = f->logits; p.labels = f->labels; p.pw = f->pixel_weight; p.cw = f->ce_class_weight;
  p.stats = reinterpret_cast<unsigned long long*>(f->stats);
  p.pb = reinterpret_cast<float*>(d->workspace);
  p.label_dtype = f->label_dtype;
  p.N = f->N; p.C = f->C; p.h = f->h; p.w = f->w; p.H = f->H; p.W = f->W;
  p.S = S;
  p.NG = 256 / S; p.GPR = S / 4; p.RT = p.NG / p.GPR;
  p.logNG = 0; while ((1 << p.logNG) < p.NG) ++p.logNG;
  p.logGPR = 0; while ((1 << p.logGPR) < p.GPR) ++p.logGPR;
  p.logRT = 0; while ((1 << p.logRT) < p.RT) ++p.logRT;
  p.ignore_index = f->ignore_index; p.acc_has_ignore = f->acc_has_ignore; p.acc_ignore = f->acc_ignore_index;
  const bool grad = d->grad_logits != nullptr || d->defer_combine;
  if (!grad) return pick_up<T, false>(p, st);
  B200SEG_REQUIRE(d->workspace != nullptr, "loss_fused: workspace is NULL");
  if (int e = pick_up<T, true>(p, st)) return e;
  if (d->defer_combine) return 0;
  return up_combine_dispatch(d->workspace, d->grad_logits, f->logit_dtype, p.N, p.C, p.h, p.w, d->grad_scale_host,
                             d->grad_out, d->use_nvalid, f->stats, st);
}

int flat_fused_dispatch(const b200seg_loss_fused_desc* d, cudaStream_t st);

int up_fused_dispatch(const b200seg_loss_fused_desc* d, cudaStream_t st) {
  const b200seg_loss_desc* f = &d->fwd;
  const bool up = (f->h != f->H) || (f->w != f->W);
  if (!up) return flat_fused_dispatch(d, st);
  int S = 0;
  B200SEG_REQUIRE(up_fast_ok(f->C, f->h, f->w, f->H, f->W, f->align_corners, &S),
                  "loss_fused: resize-fused single pass needs align_corners=False, an integer power-of-two scale in "
                  "[4,32] and C <= 32 (query b200seg_loss_fused_workspace_bytes() != 0 first)");
  switch (f->logit_dtype) {
    case B200SEG_F32: return up_run<float>(d, S, st);
    case B200SEG_BF16: return up_run<__nv_bfloat16>(d, S, st);
    case B200SEG_F16: return up_run<__half>(d, S, st);
  }
  set_error("loss_fused: unsupported logit dtype %d", f->logit_dtype);
  return 1;
}

}  // namespace b200seg
