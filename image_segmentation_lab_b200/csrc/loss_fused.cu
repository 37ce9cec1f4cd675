// Single-pass forward+backward of resize + softmax cross-entropy (+ accuracy), sm_100a.
//
// For a 'mean' / 'sum' reduction the gradient of the loss w.r.t. a pixel's logits depends on that
// pixel only (up to one global scale), so it can be produced while the logits are still in
// registers: the logits are read ONCE and the gradient written ONCE, instead of the reference's
// ~10 passes over the (N,C,H,W) up-sampled tensor (utils/ops.py:26, cross_entropy_loss.py:56-61,
// accuracy.py:41, and their autograd backwards).
//
// (b) up_fused_kernel — logits at 1/S resolution (S a power of two >= 4, align_corners=False,
//     the decode-head call models/decode_heads/decode_head.py:266-269). The (N,C,H,W) tensor is
//     never materialised in either direction.
//     Geometry: output pixel X has taps (r-1, r) with r = (X + S/2) / S; the S consecutive pixels
//     of one "run" r share their taps, likewise rows in a "band" b. A (band, run) cell of SxS
//     output pixels therefore scatters its gradient to exactly 4 low-res corners. A CTA owns one
//     band x a range of runs: thread = (row of the band, 4 consecutive pixels of one run); it
//     builds the C interpolated logits of its 4 pixels in registers (1 FFMA each from the
//     vertically-interpolated tap pair), does the soft-max, and reduces its gradient to the two
//     horizontal corners. The per-cell reduction over the S rows x S/4 threads goes through shared
//     memory in a FIXED order, and each cell's 4 corner sums are written once to a partial buffer
//     PB[n][c][band][run][2][2]; a small combine kernel adds the 4 cells around every low-res
//     logit and applies the global scale. No atomics: the backward is deterministic (ATen's
//     upsample_bilinear2d_backward is an atomicAdd scatter).
//     Bound: instruction issue (C exps + ~16 C FP32/ALU ops per output pixel), not HBM — the only
//     full-resolution tensor touched is the label map.
//
// (a) flat_fused_kernel — logits already at label resolution: the class-split register tile of
//     loss_tile.cu with the gradient store in place of the dice accumulators. HBM-bound,
//     2*C*s + L bytes per pixel instead of 3*C*s + 2L for separate forward and backward.
#include "common.cuh"

namespace b200seg {

// ================================================================================================
// (b) power-of-two up-sampling
struct UpParams {
  const void* logits;
  const void* labels;
  const float* pw;
  const float* cw;
  unsigned long long* stats;
  float* pb;        // partial buffer
  int label_dtype;
  int N, C, h, w, H, W;
  int S, logS;      // scale factor
  int NG;           // thread groups (4 px each) per row per CTA = 256 / S
  int GPR;          // groups per run = S / 4
  int RT;           // runs per tile = NG / GPR
  long long ignore_index;
  int acc_has_ignore;
  long long acc_ignore;
};

template <typename T, int CPT, bool GRAD>
__global__ void __launch_bounds__(256, (CPT <= 20 ? 2 : 1)) up_fused_kernel(const UpParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double sred[5 * 32];
  __shared__ float lam_y[32];  // cell-relative vertical weight of every row of the band
  const int C = p.C, S = p.S, NG = p.NG, GPR = p.GPR, RT = p.RT;
  const int n = blockIdx.z, b = blockIdx.y, tile = blockIdx.x;
  const int tid = threadIdx.x;
  const int i = tid / NG;        // row within the band
  const int ul = tid - i * NG;   // local group within the tile
  const int r_first = tile * RT; // first run of the tile
  const int ncol = RT + 1;

  float* patch = reinterpret_cast<float*>(smem_raw);            // [C][2][ncol]
  float2* stage = reinterpret_cast<float2*>(patch + ((C * 2 * ncol + 3) & ~3));  // [C][S][NG]

  // ---- stage the two low-res tap rows of this band (clamped) as fp32
  {
    const T* img = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * p.h * p.w;
    const int ya = b - 1 < 0 ? 0 : b - 1, yb = b > p.h - 1 ? p.h - 1 : b;
    for (int e = tid; e < C * 2 * ncol; e += 256) {
      const int k = e % ncol;
      const int rr = (e / ncol) & 1;
      const int c = e / (2 * ncol);
      int col = r_first - 1 + k;
      col = col < 0 ? 0 : (col > p.w - 1 ? p.w - 1 : col);
      patch[e] = to_float<T>(img[((size_t)c * p.h + (rr ? yb : ya)) * p.w + col]);
    }
    if (tid < S) {
      const int Y = S * b - S / 2 + tid;
      float l = 0.f;
      if (b == 0) l = 1.f;
      else if (b == p.h) l = 0.f;
      else l = ((float)Y + 0.5f) / (float)S - 0.5f - (float)(b - 1);  // exact for power-of-two S
      lam_y[tid] = (Y >= 0 && Y < p.H) ? l : -1.f;                     // -1 marks rows outside the image
    }
  }
  __syncthreads();

  const int Y = S * b - S / 2 + i;
  const int u = tile * NG + ul;
  const int X0 = 4 * u - S / 2;
  const int r = u / GPR;
  const bool row_ok = (Y >= 0 && Y < p.H);
  const bool any_ok = row_ok && r <= p.w && X0 + 3 >= 0 && X0 < p.W;
  const int k = r - r_first;
  const float ly = row_ok ? lam_y[i] : 0.f;

  float loss_acc = 0.f;
  int n_valid = 0, n_correct = 0, n_bad = 0, n_acc = 0;
  float e[CPT][4];
  float lx[4], coef[4];
  int ycl[4];
  bool pix_ok[4];

  if (any_ok) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int X = X0 + j;
      pix_ok[j] = (X >= 0 && X < p.W);
      float l;
      if (r == 0) l = 1.f;
      else if (r == p.w) l = 0.f;
      else l = ((float)X + 0.5f) / (float)S - 0.5f - (float)(r - 1);
      lx[j] = l;
    }
    // labels of the 4 pixels (X0 may be negative / past the end on the edge groups)
    long long y[4];
    const size_t lbase = ((size_t)n * p.H + Y) * p.W;
    if (X0 >= 0 && X0 + 3 < p.W && ((lbase + X0) & 3) == 0 && aligned16(p.labels)) {
      load_labels<4>(p.labels, p.label_dtype, lbase + X0, y);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = pix_ok[j] ? load_label(p.labels, p.label_dtype, lbase + X0 + j) : p.ignore_index;
    }

    float m[4];
    int idx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { m[j] = neg_inf(); idx[j] = 0; }
    const float* pc = patch + k;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      if (c < C) {
        const float a0 = pc[(c * 2 + 0) * ncol], a1 = pc[(c * 2 + 0) * ncol + 1];
        const float b0 = pc[(c * 2 + 1) * ncol], b1 = pc[(c * 2 + 1) * ncol + 1];
        const float va = fmaf(ly, b0 - a0, a0);
        const float vb = fmaf(ly, b1 - a1, a1);
        const float d = vb - va;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float z = fmaf(lx[j], d, va);
          e[c][j] = z;
          if (z > m[j]) { m[j] = z; idx[j] = c; }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) e[c][j] = neg_inf();
      }
    }
    float s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float nm = -m[j] * kLog2e;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        e[c][j] = ex2(fmaf(e[c][j], kLog2e, nm));
        acc += e[c][j];
      }
      s[j] = acc;
    }
    float pwv[4] = {1.f, 1.f, 1.f, 1.f};
    if (p.pw) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (pix_ok[j]) pwv[j] = __ldg(p.pw + lbase + X0 + j);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      coef[j] = 0.f;
      ycl[j] = -1;
      if (pix_ok[j]) {
        const long long yy = y[j];
        const bool ign = (yy == p.ignore_index);
        const bool inr = (yy >= 0 && yy < (long long)C);
        n_bad += (!ign && !inr);
        n_valid += !ign;
        if (!ign && inr) {
          const int yc = (int)yy;
          const float a0 = pc[(yc * 2 + 0) * ncol], a1 = pc[(yc * 2 + 0) * ncol + 1];
          const float b0 = pc[(yc * 2 + 1) * ncol], b1 = pc[(yc * 2 + 1) * ncol + 1];
          const float va = fmaf(ly, b0 - a0, a0), vb = fmaf(ly, b1 - a1, a1);
          const float zy = fmaf(lx[j], vb - va, va);
          const float lse = m[j] + logf(s[j]);
          const float wt = (p.cw ? __ldg(p.cw + yy) : 1.f) * pwv[j];
          loss_acc += wt * (lse - zy);
          coef[j] = wt;
          ycl[j] = yc;
        }
        const bool av = p.acc_has_ignore ? (yy != p.acc_ignore) : true;
        n_acc += av;
        n_correct += (av && (long long)idx[j] == yy);
      }
    }
    if constexpr (GRAD) {
      float rj[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) rj[j] = coef[j] / s[j];
      float2* st = stage + (size_t)i * NG + ul;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        if (c < C) {
          float gs = 0.f, gb = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float g = rj[j] * e[c][j];
            gs += g;
            gb = fmaf(lx[j], g, gb);
          }
          st[(size_t)c * S * NG] = make_float2(gs - gb, gb);
        }
      }
      // one-hot part: subtract coef at the label class (own slot: plain read-modify-write)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ycl[j] >= 0) {
          float2 v = st[(size_t)ycl[j] * S * NG];
          v.x -= (1.f - lx[j]) * coef[j];
          v.y -= lx[j] * coef[j];
          st[(size_t)ycl[j] * S * NG] = v;
        }
      }
    }
  } else if constexpr (GRAD) {
    float2* st = stage + (size_t)i * NG + ul;
    for (int c = 0; c < C; ++c) st[(size_t)c * S * NG] = make_float2(0.f, 0.f);
  }

  if constexpr (GRAD) {
    __syncthreads();
    // fixed-order per-cell reduction: item = (class, run of the tile, corner row)
    const int items = C * RT * 2;
    for (int it = tid; it < items; it += 256) {
      const int cr = it & 1;
      const int rl = (it >> 1) % RT;
      const int c = (it >> 1) / RT;
      const int rr = r_first + rl;
      if (rr > p.w) continue;
      float sa = 0.f, sb = 0.f;
      for (int ii = 0; ii < S; ++ii) {
        const float l = lam_y[ii];
        if (l < 0.f) continue;
        const float wy = cr ? l : 1.f - l;
        const float2* row = stage + ((size_t)c * S + ii) * NG + rl * GPR;
        float ra = 0.f, rb = 0.f;
        for (int q = 0; q < GPR; ++q) { ra += row[q].x; rb += row[q].y; }
        sa = fmaf(wy, ra, sa);
        sb = fmaf(wy, rb, sb);
      }
      float2* dst = reinterpret_cast<float2*>(p.pb) +
                    ((((size_t)n * C + c) * (p.h + 1) + b) * (p.w + 1) + rr) * 2 + cr;
      *dst = make_float2(sa, sb);
    }
  }

  double red[5] = {(double)loss_acc, (double)n_valid, (double)n_correct, (double)n_bad, (double)n_acc};
  block_sum<double, 5>(red, sred);
  if (tid == 0) {
    atomicAdd(reinterpret_cast<double*>(p.stats + B200SEG_ST_CE_SUM), red[0]);
    atomicAdd(p.stats + B200SEG_ST_N_VALID, (unsigned long long)red[1]);
    atomicAdd(p.stats + B200SEG_ST_N_CORRECT, (unsigned long long)red[2]);
    if (red[3] != 0.0) atomicAdd(p.stats + B200SEG_ST_N_BAD, (unsigned long long)red[3]);
    atomicAdd(p.stats + B200SEG_ST_N_ACC, (unsigned long long)red[4]);
  }
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

template <typename T, int CPT, bool GRAD> static int launch_up(const UpParams& p, cudaStream_t st) {
  const int ncol = p.RT + 1;
  size_t smem = (size_t)((p.C * 2 * ncol + 3) & ~3) * 4;
  if (GRAD) smem += (size_t)p.C * 256 * sizeof(float2);
  auto k = up_fused_kernel<T, CPT, GRAD>;
  static bool attr = false;
  if (!attr) {
    B200SEG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  B200SEG_REQUIRE(smem <= 100 * 1024, "loss_fused: shared-memory tile too large (%zu bytes)", smem);
  const int NU = (p.W + p.S / 2 + 3) / 4;
  dim3 grid((NU + p.NG - 1) / p.NG, p.h + 1, p.N);
  k<<<grid, 256, smem, st>>>(p);
  count_launch();
  return check_launch("up_fused_kernel");
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

template <typename T>
__global__ void __launch_bounds__(256) scale_inplace_kernel(T* __restrict__ x, long long n, const float* __restrict__ g) {
  const float s = __ldg(g);
  if (s == 1.f) return;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = from_float<T>(to_float<T>(x[i]) * s);
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
  p.S = S; p.logS = 0;
  while ((1 << p.logS) < S) ++p.logS;
  p.NG = 256 / S; p.GPR = S / 4; p.RT = p.NG / p.GPR;
  p.ignore_index = f->ignore_index; p.acc_has_ignore = f->acc_has_ignore; p.acc_ignore = f->acc_ignore_index;
  const bool grad = d->grad_logits != nullptr || d->defer_combine;
  if (!grad) return pick_up<T, false>(p, st);
  B200SEG_REQUIRE(d->workspace != nullptr, "loss_fused: workspace is NULL");
  if (int e = pick_up<T, true>(p, st)) return e;
  if (d->defer_combine) return 0;
  return up_combine_dispatch(d->workspace, d->grad_logits, f->logit_dtype, p.N, p.C, p.h, p.w, d->grad_scale_host,
                             d->grad_out, d->use_nvalid, f->stats, st);
}

// ================================================================================================
// (a) logits at label resolution: class-split register tile, gradient written in the same pass
struct FlatParams {
  const void* logits;
  const void* labels;
  const float* pw;
  const float* cw;
  const float* grad_out;
  unsigned long long* stats;
  void* grad;
  float scale_host;
  int label_dtype;
  int N, C;
  long long HW;
  long long ignore_index;
  int acc_has_ignore;
  long long acc_ignore;
  int G, cpg, PG, tiles;
};

__device__ __forceinline__ void fgroup_barrier(int pg, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(pg + 1), "r"(nthreads) : "memory");
}

template <typename T, int V, int CPT>
__global__ void __launch_bounds__(512) flat_fused_kernel(const FlatParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double sred[5 * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = p.G, PG = p.PG, C = p.C;
  const int pg = warp / G, g = warp - pg * G;
  const int c0 = g * p.cpg;
  const int c1 = (c0 + p.cpg < C) ? c0 + p.cpg : C;
  const int n = blockIdx.y;
  const long long HW = p.HW;
  constexpr int PXW = 32 * V;

  float* exch_m = reinterpret_cast<float*>(smem_raw);            // [PG][G][PXW]
  int* exch_i = reinterpret_cast<int*>(exch_m + PG * G * PXW);
  float* exch_s = reinterpret_cast<float*>(exch_i + PG * G * PXW);
  float* exch_k = exch_s + PG * G * PXW;                          // [PG][PXW]
  int* exch_y = reinterpret_cast<int*>(exch_k + PG * PXW);        // [PG][PXW]

  const float Gs = p.scale_host * (p.grad_out ? __ldg(p.grad_out) : 1.f);
  const T* img = reinterpret_cast<const T*>(p.logits) + (size_t)n * C * HW;
  T* gimg = reinterpret_cast<T*>(p.grad) + (size_t)n * C * HW;

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
    long long y[V];
    float pwv[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { y[v] = p.ignore_index; pwv[v] = 1.f; }
    if (g == 0 && active) {
      load_labels<V>(p.labels, p.label_dtype, (size_t)n * HW + px0, y);
      if (p.pw) load_vec<float, V>(p.pw + (size_t)n * HW + px0, pwv);
    }
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
    const int pslot = pg * PXW + lane * V;
    float kk[V];
    int yc[V];
    if (g == 0) {  // per-pixel gradient coefficient and label, shared with the other class groups
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const long long yy = y[v];
        const bool valid = active && (yy != p.ignore_index) && yy >= 0 && yy < (long long)C;
        kk[v] = valid ? pwv[v] * (p.cw ? __ldg(p.cw + yy) : 1.f) : 0.f;
        yc[v] = valid ? (int)yy : -1;
      }
    }
    if (G > 1) {
      const int slot = (pg * G + g) * PXW + lane * V;
#pragma unroll
      for (int v = 0; v < V; ++v) { exch_m[slot + v] = m[v]; exch_i[slot + v] = idx[v]; }
      if (g == 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) { exch_k[pslot + v] = kk[v]; exch_y[pslot + v] = yc[v]; }
      }
      fgroup_barrier(pg, 32 * G);
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
        kk[v] = exch_k[pslot + v];
        yc[v] = exch_y[pslot + v];
      }
    }
    float s[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float nm = active ? -m[v] * kLog2e : 0.f;
      float ls = 0.f;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        z[i][v] = ex2(fmaf(z[i][v], kLog2e, nm));
        ls += z[i][v];
      }
      s[v] = ls;
    }
    if (G > 1) {
      const int slot = (pg * G + g) * PXW + lane * V;
#pragma unroll
      for (int v = 0; v < V; ++v) exch_s[slot + v] = s[v];
      fgroup_barrier(pg, 32 * G);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float t = 0.f;
        for (int gg = 0; gg < G; ++gg) t += exch_s[(pg * G + gg) * PXW + lane * V + v];
        s[v] = t;
      }
    }
    if (active) {
      float rr[V], kg[V];
#pragma unroll
      for (int v = 0; v < V; ++v) { kg[v] = kk[v] * Gs; rr[v] = kg[v] / s[v]; }
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        if (c0 + i < c1) {
          float gr[V];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            gr[v] = rr[v] * z[i][v];
            if (c0 + i == yc[v]) gr[v] -= kg[v];
          }
          store_vec<T, V>(gimg + (size_t)(c0 + i) * HW + px0, gr);
        }
      }
      if (g == 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const long long yy = y[v];
          const bool ign = (yy == p.ignore_index);
          const bool inr = (yy >= 0 && yy < (long long)C);
          n_bad += (!ign && !inr);
          n_valid += !ign;
          if (!ign && inr) {
            const float zy = to_float<T>(img[(size_t)yy * HW + px0 + v]);
            loss_acc += kk[v] * (m[v] + logf(s[v]) - zy);
          }
          const bool av = p.acc_has_ignore ? (yy != p.acc_ignore) : true;
          n_acc += av;
          n_correct += (av && (long long)idx[v] == yy);
        }
      }
    }
  }
  double r[5] = {(double)loss_acc, (double)n_valid, (double)n_correct, (double)n_bad, (double)n_acc};
  block_sum<double, 5>(r, sred);
  if (threadIdx.x == 0) {
    atomicAdd(reinterpret_cast<double*>(p.stats + B200SEG_ST_CE_SUM), r[0]);
    atomicAdd(p.stats + B200SEG_ST_N_VALID, (unsigned long long)r[1]);
    atomicAdd(p.stats + B200SEG_ST_N_CORRECT, (unsigned long long)r[2]);
    if (r[3] != 0.0) atomicAdd(p.stats + B200SEG_ST_N_BAD, (unsigned long long)r[3]);
    atomicAdd(p.stats + B200SEG_ST_N_ACC, (unsigned long long)r[4]);
  }
}

template <typename T, int V, int CPT> static int launch_flat(FlatParams p, cudaStream_t st) {
  p.G = (p.C + CPT - 1) / CPT;
  p.cpg = (p.C + p.G - 1) / p.G;
  p.PG = 4 / p.G;
  if (p.PG < 1) p.PG = 1;
  const long long per_tile = (long long)p.PG * 32 * V;
  p.tiles = (int)((p.HW + per_tile - 1) / per_tile);
  const size_t smem = (size_t)3 * p.PG * p.G * 32 * V * 4 + (size_t)2 * p.PG * 32 * V * 4;
  int gx = (kSMs * 8 + p.N - 1) / p.N;
  if (gx > p.tiles) gx = p.tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, p.N);
  flat_fused_kernel<T, V, CPT><<<grid, 32 * p.G * p.PG, smem, st>>>(p);
  count_launch();
  return check_launch("flat_fused_kernel");
}

template <typename T> static int flat_run(const b200seg_loss_fused_desc* d, cudaStream_t st) {
  const b200seg_loss_desc* f = &d->fwd;
  B200SEG_REQUIRE(!d->use_nvalid, "loss_fused: avg_non_ignore needs the two-pass path at label resolution");
  B200SEG_REQUIRE(f->C <= 512, "loss_fused: at most 512 classes (got %d)", f->C);
  FlatParams p = {};
  p.logits = f->logits; p.labels = f->labels; p.pw = f->pixel_weight; p.cw = f->ce_class_weight;
  p.grad_out = d->grad_out; p.stats = reinterpret_cast<unsigned long long*>(f->stats); p.grad = d->grad_logits;
  p.scale_host = d->grad_scale_host; p.label_dtype = f->label_dtype;
  p.N = f->N; p.C = f->C; p.HW = (long long)f->H * f->W;
  p.ignore_index = f->ignore_index; p.acc_has_ignore = f->acc_has_ignore; p.acc_ignore = f->acc_ignore_index;
  const bool vec = (p.HW % 2 == 0) && aligned16(f->logits) && aligned16(f->labels) && aligned16(d->grad_logits) &&
                   (!p.pw || aligned16(p.pw));
  const int C = p.C;
  const int cpt = C <= 8 ? 8 : (C <= 16 ? 16 : (C <= 32 ? 32 : (C <= 256 ? 16 : 32)));
  if (vec) {
    if (cpt == 8) return launch_flat<T, 2, 8>(p, st);
    if (cpt == 16) return launch_flat<T, 2, 16>(p, st);
    return launch_flat<T, 2, 32>(p, st);
  }
  if (cpt == 8) return launch_flat<T, 1, 8>(p, st);
  if (cpt == 16) return launch_flat<T, 1, 16>(p, st);
  return launch_flat<T, 1, 32>(p, st);
}

int up_fused_dispatch(const b200seg_loss_fused_desc* d, cudaStream_t st) {
  const b200seg_loss_desc* f = &d->fwd;
  const bool up = (f->h != f->H) || (f->w != f->W);
  if (!up) {
    B200SEG_REQUIRE(d->grad_logits != nullptr, "loss_fused: grad_logits is NULL");
    switch (f->logit_dtype) {
      case B200SEG_F32: return flat_run<float>(d, st);
      case B200SEG_BF16: return flat_run<__nv_bfloat16>(d, st);
      case B200SEG_F16: return flat_run<__half>(d, st);
    }
    set_error("loss_fused: unsupported logit dtype %d", f->logit_dtype);
    return 1;
  }
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
