// CE + Dice (+ top-1 accuracy) for MANY classes (32 < C <= 152) with ONE read of the logits per direction, sm_100a.
//
// Replaces, for class counts that do not fit one warp's register tile (ADE20K: 150 classes):
//   forward   F.cross_entropy + weight_reduce_loss + accuracy + DiceLoss.forward
//             (models/losses/cross_entropy_loss.py:56-72, models/losses/utils.py:48-80, models/losses/accuracy.py:41-60,
//              models/losses/dice_loss.py:23-58,103-134: F.softmax, the int64 (N,H,W,C) one-hot, the C-iteration loop)
//   backward  the autograd graph behind them
// by ONE pass over the logits each way. The streaming kernels this supersedes (loss_stream.cu + loss_dice.cu) needed two
// reads forward (log-sum-exp, then sum_px p_c^2) and two reads + one write backward (sum_c p_c g_c, then the gradient),
// because a pixel's 150 soft-max values do not fit one thread and a register tile has no loads in flight while it is
// computed on.
//
// Structure — a "class-sliced" bulk-copy pipeline:
//   * A persistent CTA owns a contiguous range of 128-pixel tiles. A producer warp issues TENSOR-MAP TMA copies
//     (`cp.async.bulk.tensor.2d`, global -> shared, completion on an mbarrier): one instruction per box of
//     (128 bytes of pixels) x (all C class rows), i.e. 2 (16-bit) or 4 (fp32) per tile, plus one plain bulk copy for the
//     label row (and the saved log-sum-exp row in the backward), 2-5 stages ahead: bytes in flight are set by the stage
//     count, not by registers. (One `cp.async.bulk` per 256-byte class row — the first version — was bound by the copy
//     ISSUE rate: ~70 cycles per copy per SM, 1.2 ms per pass at ADE20K shape.)
//   * Boxes land 128B-swizzled (16-byte chunk index XOR row index mod 8). Consumer warps (8 per tile): a warp owns 16
//     pixels of the tile, and inside the warp the 32 lanes form 4 pixel groups x 8 CLASS SLICES: lane (j, g) holds
//     classes {8 i + j} of its group's pixels, read with one 4-byte (16-bit logits, 2 pixels) or 16-byte (fp32, 4 pixels)
//     shared-memory access per class row; row 8 i + j sits at chunk (c XOR j), so the 8 slices x 4 groups of a wavefront
//     hit 32 distinct banks.
//   * Per-pixel reductions over the class dimension (max, sum of exponentials, the backward's dot product) are CPT
//     register operations plus THREE xor-shuffles — the slices of a pixel live in one warp, so there is no CTA barrier
//     and no shared-memory exchange. The maximum is taken on the PACKED 16-bit pairs (HMNMX2).
//   * One MUFU.EX2 per element per direction: the exponentials stay in registers between the sum and the p^2 / gradient
//     sweep.
//   * Per-pixel scalar work (label decode, label logit, log, loss, accuracy, one-hot Dice terms) is done once per pixel
//     by lane (j = pixel-in-group, g); its contribution to the backward's dot product rides on the same shuffle tree.
//   * Backward: the tile is overwritten in place with the gradient and handed to a store warp (`cp.async.bulk`
//     shared -> global).
// Tiles are walked back to front by the backward kernel: with the same tile partition as the forward, every CTA starts on
// the part of its range the forward left in the 126 MB L2.
//
// Top-1 accuracy: the label's class counts as the arg-max iff its logit equals the pixel's maximum (an exact tie with
// another class counts as correct; torch.topk's choice among ties is unspecified) — same rule as the resize-fused kernel.
//
// Bound: HBM (forward is close to issue balance at 16-bit width: ~8.5 thread instructions per element).
// Algorithmic bytes per launch: forward N*C*H*W*s + N*H*W*(L+4), backward 2*N*C*H*W*s + N*H*W*(L+4).
#include <cuda.h>   // CUtensorMap (the encode entry point is fetched through the runtime: no link against libcuda)

#include "bulk_pipe.cuh"
#include "common.cuh"

namespace b200seg {

constexpr int kCsSlices = 8;      // class slices per pixel group (lane bits 2..4)
constexpr int kCsLanePx = 4;      // consecutive pixels per lane
constexpr int kCsWarpPx = 16;     // pixels per consumer warp per tile (4 groups x 4 pixels)
constexpr int kCsGroupWarps = 8;  // consumer warps per tile (a "group"); a CTA runs G groups on alternating tiles
constexpr int kCsMaxStages = 6;
constexpr int kCsBoxBytes = 128;   // inner extent of a TMA box = one swizzle row
constexpr int kCsMaxClasses = 152;          // 8 slices x 19 classes
constexpr int kCsBig = (int)0x80000000;     // a label that does not fit int32: never a class, never ignored
constexpr int kCsNever = (int)0x80000001;   // an ignore value no decoded label can take
__device__ __forceinline__ int cs_label32(const unsigned char* row, int dt, int t) {
  const long long v = smem_label(row, dt, t);
  return (v == (long long)(int)v && (int)v != kCsNever) ? (int)v : kCsBig;
}

struct CsParams {
  const void* logits;
  const void* labels;
  const float* pw;
  const float* cw;
  float* lse;                        // forward: written; backward: read (bulk-copied with the tile)
  unsigned long long* stats;
  double* dice_part;                 // forward: (N,C,3) [sum p_y v, sum p^2, sum t]
  const float* dice_coef;            // backward: (N,C,2) [alpha, beta]
  const float* dice_grad_out;
  const float* ce_grad_out;
  void* grad;
  int label_dtype, label_bytes;
  int N, C;
  long long HW;
  int tiles_per_image;
  long long total_tiles, tiles_per_cta;
  int stages, stage_bytes, block_bytes, label_off, lse_off;   // block = one TMA box in shared memory: CP rows x 128 bytes
  int flags;
  int ignore32, dice_ignore32;       // ignore values squashed to int32 (kCsNever when they do not fit: never matches)
  int acc_has_ignore;
  int acc_ignore32;
  float ce_scale_host;
  int ce_use_nvalid;
};

// PX consecutive pixels of one class row in the storage format (16-bit logits stay packed): a lane's 4 pixels of a
// tile are processed as 4 / PX passes — two passes of 2 pixels for 16-bit logits (the live exponentials of 4 pixels x 19
// classes do not fit the 96 registers two resident CTAs leave a thread), one pass of 4 for fp32 (one CTA per SM).
template <typename T> struct CsCfg {
  static constexpr int PX = sizeof(T) == 2 ? 2 : 4;
  static constexpr int kPasses = kCsLanePx / PX;
  static constexpr int kWords = PX * (int)sizeof(T) / 4;      // 1 (16-bit) or 4 (fp32)
  static constexpr int kBoxPx = kCsBoxBytes / (int)sizeof(T);  // pixels per TMA box row: 64 (16-bit) or 32 (fp32)
};
template <typename T> struct CsRow {
  uint32_t w[CsCfg<T>::kWords];
};
template <typename T> __device__ __forceinline__ CsRow<T> cs_load(const unsigned char* p) {
  CsRow<T> r;
  if constexpr (sizeof(T) == 4) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    r.w[0] = t.x; r.w[1] = t.y; r.w[2] = t.z; r.w[3] = t.w;
  } else {
    r.w[0] = *reinterpret_cast<const uint32_t*>(p);
  }
  return r;
}
template <typename T> __device__ __forceinline__ void cs_unpack(const CsRow<T>& r, float (&z)[CsCfg<T>::PX]) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) z[k] = __uint_as_float(r.w[k]);
  } else {
    unpack2<T>(r.w[0], z[0], z[1]);
  }
}
template <typename T> __device__ __forceinline__ uint32_t max2_packed(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t max2_packed<__nv_bfloat16>(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t max2_packed<__half>(uint32_t a, uint32_t b) {
  const __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <typename T> __device__ __forceinline__ void cs_max(CsRow<T>& m, const CsRow<T>& z) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) m.w[k] = __float_as_uint(fmaxf(__uint_as_float(m.w[k]), __uint_as_float(z.w[k])));
  } else {
    m.w[0] = max2_packed<T>(m.w[0], z.w[0]);
  }
}
template <typename T> __device__ __forceinline__ void cs_store(unsigned char* p, const float (&g)[CsCfg<T>::PX]) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<uint4*>(p) = make_uint4(__float_as_uint(g[0]), __float_as_uint(g[1]), __float_as_uint(g[2]), __float_as_uint(g[3]));
  } else {
    *reinterpret_cast<uint32_t*>(p) = pack2<T>(g[0], g[1]);
  }
}
template <typename T> __device__ __forceinline__ uint32_t neg_inf_word() {
  if constexpr (sizeof(T) == 4) return 0xff800000u;
  else if constexpr (Elem<T>::kDtype == B200SEG_BF16) return 0xff80ff80u;
  else return 0xfc00fc00u;
}
__device__ __forceinline__ float pick4(const float (&a)[4], int k) {
  return k == 0 ? a[0] : (k == 1 ? a[1] : (k == 2 ? a[2] : a[3]));
}

// One tile of the CTA's range: image, first pixel, pixel count.
struct CsTile {
  int n, npx;
  long long px0;
};
// Walks a CTA's tile range forward (or backward) in steps of `step` tiles without a division per tile.
struct CsWalker {
  int n, tin, tpi, step, TP;
  long long HW;
  __device__ __forceinline__ void init(const CsParams& p, long long t_first, int step_, int TP_) {
    tpi = p.tiles_per_image; step = step_; TP = TP_; HW = p.HW;
    n = (int)(t_first / tpi);
    tin = (int)(t_first - (long long)n * tpi);
  }
  __device__ __forceinline__ CsTile tile() const {
    CsTile r;
    r.n = n;
    r.px0 = (long long)tin * TP;
    r.npx = (int)((HW - r.px0 < TP) ? HW - r.px0 : TP);
    return r;
  }
  __device__ __forceinline__ void forward() {
    tin += step;
    while (tin >= tpi) { tin -= tpi; ++n; }
  }
  __device__ __forceinline__ void backward() {
    tin -= step;
    while (tin < 0) { tin += tpi; --n; }
  }
};

// Shared-memory address of (class row, pixel) inside a stage: box = pixel's 128-byte column block, 128B swizzle inside.
template <typename T> __device__ __forceinline__ unsigned cs_offset(const CsParams& p, int row, int px) {
  const unsigned byte = (unsigned)px * (unsigned)sizeof(T);
  const unsigned blk = byte >> 7, inb = byte & 127u;
  return blk * (unsigned)p.block_bytes + (unsigned)row * 128u + ((((inb >> 4) ^ ((unsigned)row & 7u)) << 4) | (inb & 15u));
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<unsigned long long>(map)),
               "r"(c0), "r"(c1), "r"(smem_u32(smem_src))
               : "memory");
}

// Producer warp body shared by both directions (REVERSE walks the range back to front): lane 0 issues the boxes of a tile.
template <typename T, bool WITH_LSE, bool REVERSE>
__device__ __forceinline__ void cs_producer(const CsParams& p, const CUtensorMap* map, unsigned char* smem,
                                            unsigned long long* full_bar, unsigned long long* empty_bar, long long t0,
                                            long long t1, int TP, int lane) {
  constexpr int BOXPX = CsCfg<T>::kBoxPx;
  const int C = p.C, NS = p.stages;
  CsWalker wk;
  wk.init(p, REVERSE ? t1 - 1 : t0, 1, TP);
  const int nt = (int)(t1 - t0);
  StageRing ring;
  ring.init(0, NS, 1);
  for (int k = 0; k < nt; ++k, ring.advance()) {
    const int s = ring.s;
    if (k >= NS) mbar_wait(&empty_bar[s], ring.ph ^ 1);
    if (lane == 0) {
      const CsTile tl = wk.tile();
      const int nbox = (tl.npx + BOXPX - 1) / BOXPX;           // boxes that start inside the image (a partial one is zero-filled)
      const unsigned lab_bytes = (unsigned)(tl.npx * p.label_bytes);
      const unsigned lse_bytes = WITH_LSE ? (unsigned)(tl.npx * 4) : 0u;
      unsigned char* stage = smem + (size_t)s * p.stage_bytes;
      mbar_arrive_expect_tx(&full_bar[s], (unsigned)nbox * (unsigned)C * (unsigned)kCsBoxBytes + lab_bytes + lse_bytes);
      for (int bx = 0; bx < nbox; ++bx)
        tma_load_2d(stage + (size_t)bx * p.block_bytes, map, (int)(tl.px0 + (long long)bx * BOXPX), tl.n * C, &full_bar[s]);
      bulk_g2s(stage + p.label_off, reinterpret_cast<const char*>(p.labels) + ((size_t)tl.n * p.HW + tl.px0) * p.label_bytes,
               lab_bytes, &full_bar[s]);
      if (WITH_LSE)
        bulk_g2s(stage + p.lse_off, reinterpret_cast<const char*>(p.lse + (size_t)tl.n * p.HW + tl.px0), lse_bytes, &full_bar[s]);
    }
    if (REVERSE) wk.backward(); else wk.forward();
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <typename T, int CPT, int G>
__global__ void __launch_bounds__((kCsGroupWarps * G + 1) * 32, 1) cs_fwd_kernel(const CsParams p, const __grid_constant__ CUtensorMap tmap) {
  constexpr int NWG = kCsGroupWarps, NW = NWG * G;
  constexpr int TP = NWG * kCsWarpPx;
  constexpr int CP = CPT * kCsSlices;
  extern __shared__ unsigned char smem_dyn[];
  // stages are 1024-byte aligned: the 128B swizzle is a function of the shared-memory ADDRESS bits
  unsigned char* const smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ __align__(8) unsigned long long full_bar[kCsMaxStages], empty_bar[kCsMaxStages];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, NS = p.stages;
  const bool dice = (p.flags & B200SEG_WANT_DICE) != 0;
  float* bins = reinterpret_cast<float*>(smem_raw + (size_t)NS * p.stage_bytes);   // [NW][2][C]
  __shared__ float cw_s[kCsMaxClasses];                                           // CE class weights (1 when absent)
  for (int c = tid; c < C; c += blockDim.x) cw_s[c] = p.cw ? __ldg(p.cw + c) : 1.f;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NWG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // rows C..CP-1 of every stage are never written by the bulk copies: -inf once (exp -> 0, max unaffected)
  {
    constexpr int kBlocks = TP * (int)sizeof(T) / kCsBoxBytes;     // TMA boxes per tile
    const int pad_words = (CP - C) * (kCsBoxBytes / 4);            // per block: rows C..CP-1 (swizzling permutes inside a row)
    for (int s = 0; s < NS; ++s)
      for (int bx = 0; bx < kBlocks; ++bx)
        for (int i = tid; i < pad_words; i += blockDim.x)
          reinterpret_cast<uint32_t*>(smem_raw + (size_t)s * p.stage_bytes + (size_t)bx * p.block_bytes + (size_t)C * kCsBoxBytes)[i] =
              neg_inf_word<T>();
    for (int i = tid; i < NW * 2 * C; i += blockDim.x) bins[i] = 0.f;
  }
  __syncthreads();

  const long long t0 = (long long)blockIdx.x * p.tiles_per_cta;
  const long long t1 = (t0 + p.tiles_per_cta < p.total_tiles) ? t0 + p.tiles_per_cta : p.total_tiles;
  float loss_acc = 0.f;
  int n_valid = 0, n_correct = 0, n_bad = 0, n_acc = 0;

  if (warp == NW) {
    cs_producer<T, false, false>(p, &tmap, smem_raw, full_bar, empty_bar, t0, t1, TP, lane);
  } else {
    constexpr int PX = CsCfg<T>::PX, kPasses = CsCfg<T>::kPasses;
    const int j = lane & 7, g = lane >> 3;                   // class slice / pixel group
    const int grp = warp / NWG;
    const int pxo = (warp % NWG) * kCsWarpPx;                // first pixel of this warp inside the tile
    float* A_w = bins + (size_t)warp * 2 * C;
    float* T_w = A_w + C;
    float acc[CPT];                                          // sum_px p_c^2 of this lane's classes for the current image
#pragma unroll
    for (int i = 0; i < CPT; ++i) acc[i] = 0.f;
    int n_cur = -1;

    auto flush_image = [&](int n) {
      if (dice) {
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          float t = acc[i];
          t += __shfl_xor_sync(0xffffffffu, t, 8);
          t += __shfl_xor_sync(0xffffffffu, t, 16);
          const int c = i * kCsSlices + j;
          if (g == 0 && c < C) atomicAdd(p.dice_part + ((size_t)n * C + c) * 3 + 1, (double)t);
          acc[i] = 0.f;
        }
        __syncwarp();
        for (int c = lane; c < C; c += 32) {
          const float a = A_w[c], t = T_w[c];
          if (t != 0.f) {
            atomicAdd(p.dice_part + ((size_t)n * C + c) * 3 + 0, (double)a);
            atomicAdd(p.dice_part + ((size_t)n * C + c) * 3 + 2, (double)t);
            A_w[c] = 0.f;
            T_w[c] = 0.f;
          }
        }
        __syncwarp();
      }
    };

    CsWalker wk;
    wk.init(p, t0 + grp, G, TP);
    StageRing ring;
    ring.init(grp, NS, G);
    for (int k = grp; k < (int)(t1 - t0); k += G, wk.forward(), ring.advance()) {   // group grp consumes tiles grp, grp + G, ...
      const int s = ring.s;
      const CsTile tl = wk.tile();
      if (tl.n != n_cur) {
        if (n_cur >= 0) flush_image(n_cur);
        n_cur = tl.n;
      }
      mbar_wait(&full_bar[s], ring.ph);
      const unsigned char* stage = smem_raw + (size_t)s * p.stage_bytes;
      float m[4], S[4];                                     // per pixel of this lane (all 4, both passes)
#pragma unroll
      for (int h = 0; h < kPasses; ++h) {
        const int pxh = pxo + h * 4 * PX + g * PX;          // first pixel of this lane in pass h
        const unsigned char* col = stage + cs_offset<T>(p, j, pxh);   // class row 8 i + j: col + i * 1024 (same swizzle phase)
        const bool in = pxh < tl.npx;   // npx is a multiple of 16 bytes / sizeof(T) >= PX: a lane's pixels are all in or all out

        // sweep 1: the lane's CPT x PX logits (kept packed) and their maximum per pixel
        CsRow<T> zr[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) zr[i] = cs_load<T>(col + (size_t)i * (kCsSlices * kCsBoxBytes));
        CsRow<T> mx = zr[0];
#pragma unroll
        for (int i = 1; i < CPT; ++i) cs_max<T>(mx, zr[i]);
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
          CsRow<T> ot;
#pragma unroll
          for (int q = 0; q < CsCfg<T>::kWords; ++q) ot.w[q] = __shfl_xor_sync(0xffffffffu, mx.w[q], o);
          cs_max<T>(mx, ot);
        }
        float mh[PX], Sh[PX];
        cs_unpack<T>(mx, mh);
        // sweeps 2 and 3 work on PIXEL PAIRS with the packed fp32 forms (FFMA2 / FADD2 / FMUL2: one issue slot for two
        // pixels; same rounding as the scalar forms) — the forward is issue bound
        constexpr int PP = PX / 2;
        float2 nm2[PP], Sa2[PP], Sb2[PP];
        const float2 l2e = make_float2(kLog2e, kLog2e);
#pragma unroll
        for (int q = 0; q < PP; ++q) {
          nm2[q] = make_float2(-mh[2 * q] * kLog2e, -mh[2 * q + 1] * kLog2e);
          Sa2[q] = make_float2(0.f, 0.f);
          Sb2[q] = make_float2(0.f, 0.f);
        }

        // sweep 2: exponentials (kept) and their sum (two partial sums per pixel: shorter dependent chains)
        float2 e2[CPT][PP];
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          float z[PX];
          cs_unpack<T>(zr[i], z);
#pragma unroll
          for (int q = 0; q < PP; ++q) {
            const float2 x = __ffma2_rn(make_float2(z[2 * q], z[2 * q + 1]), l2e, nm2[q]);
            e2[i][q] = make_float2(ex2(x.x), ex2(x.y));
            if (i & 1) Sb2[q] = __fadd2_rn(Sb2[q], e2[i][q]);
            else Sa2[q] = __fadd2_rn(Sa2[q], e2[i][q]);
          }
        }
#pragma unroll
        for (int q = 0; q < PP; ++q) {
          const float2 t = __fadd2_rn(Sa2[q], Sb2[q]);
          Sh[2 * q] = t.x;
          Sh[2 * q + 1] = t.y;
        }
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
          for (int v = 0; v < PX; ++v) Sh[v] += __shfl_xor_sync(0xffffffffu, Sh[v], o);
        }

        // sweep 3: sum_px p^2 per class (Dice denominator: NOT masked by the valid mask, dice_loss.py:55-56)
        if (dice && in) {
          float2 r2[PP];
#pragma unroll
          for (int q = 0; q < PP; ++q) r2[q] = make_float2(fast_rcp(Sh[2 * q]), fast_rcp(Sh[2 * q + 1]));
#pragma unroll
          for (int i = 0; i < CPT; ++i) {
#pragma unroll
            for (int q = 0; q < PP; ++q) {
              const float2 pr = __fmul2_rn(e2[i][q], r2[q]);
              acc[i] = fmaf(pr.y, pr.y, fmaf(pr.x, pr.x, acc[i]));
            }
          }
        }
#pragma unroll
        for (int v = 0; v < PX; ++v) { m[h * PX + v] = mh[v]; S[h * PX + v] = Sh[v]; }
      }

      // per-pixel scalar work: lane (j < 4, g) owns pixel (pass j / PX, v = j % PX) of group g
      int cls = -1;
      float pyv = 0.f;
      const int t_px = pxo + (j / PX) * 4 * PX + g * PX + (j % PX);
      if (j < 4 && t_px < tl.npx) {
        const float m_own = pick4(m, j), S_own = pick4(S, j);
        const float lse = m_own + fast_log(S_own);
        const int yy = cs_label32(stage + p.label_off, p.label_dtype, t_px);
        const bool ign = (yy == p.ignore32);
        const bool inr = (unsigned)yy < (unsigned)C;
        const bool valid = !ign && inr;
        n_bad += (!ign && !inr);
        n_valid += !ign;
        const int ycc = yy < 0 ? 0 : (yy >= C ? C - 1 : yy);
        const float zy = to_float<T>(*reinterpret_cast<const T*>(stage + cs_offset<T>(p, ycc, t_px)));
        const size_t gpx = (size_t)tl.n * p.HW + tl.px0 + t_px;
        if (valid && (p.flags & B200SEG_WANT_CE)) {
          const float wt = cw_s[ycc];
          const float pwv = p.pw ? __ldg(p.pw + gpx) : 1.f;
          loss_acc = fmaf(wt * pwv, lse - zy, loss_acc);
        }
        const bool av = p.acc_has_ignore ? (yy != p.acc_ignore32) : true;
        n_acc += av;
        n_correct += (av && inr && zy == m_own);
        if (dice) {
          const bool dv = (yy != p.dice_ignore32);                     // valid_mask
          cls = ycc;                                                   // one-hot of the CLAMPED label (dice_loss.py:119-122)
          pyv = dv ? ex2((zy - lse) * kLog2e) : 0.f;
        }
        if (p.lse) p.lse[gpx] = lse;
      }
      if (dice) onehot_bins_add(A_w, T_w, cls, pyv, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    if (n_cur >= 0) flush_image(n_cur);
  }
  cta_flush_stats(loss_acc, n_valid, n_correct, n_bad, n_acc, p.stats, (p.flags & B200SEG_WANT_CE) != 0);
}

// ------------------------------------------------------------------------------------------------ backward
// grad_c = p_c (b_c p_c + sub) - onehot_c (p_y da + kk),   sub = kk - dot,   dot = sum_c b_c p_c^2 - da p_y
//   b_c = 2 god beta[n][c],  da = god alpha[n][y] (Dice-valid pixels),  kk = Gce pw cw[y] (CE-valid pixels)
template <typename T, int CPT, int G>
__global__ void __launch_bounds__((kCsGroupWarps * G + 2) * 32, 1) cs_bwd_kernel(const CsParams p, const __grid_constant__ CUtensorMap tmap,
                                                                                  const __grid_constant__ CUtensorMap tmap_grad) {
  constexpr int NWG = kCsGroupWarps, NW = NWG * G;
  constexpr int TP = NWG * kCsWarpPx;
  extern __shared__ unsigned char smem_dyn[];
  // stages are 1024-byte aligned: the 128B swizzle is a function of the shared-memory ADDRESS bits
  unsigned char* const smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ __align__(8) unsigned long long full_bar[kCsMaxStages], done_bar[kCsMaxStages], empty_bar[kCsMaxStages];
  // per-image Dice coefficients of the image a consumer group is working on ([alpha | 2 beta], scaled by the upstream
  // gradient) and the CE class weights: the per-pixel scalar work reads them from shared memory, not through L2
  __shared__ float coef_s[G][2][kCsMaxClasses];
  __shared__ float cw_s[kCsMaxClasses];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, NS = p.stages;
  constexpr int CP = CPT * kCsSlices;
  for (int c = tid; c < C; c += blockDim.x) cw_s[c] = p.cw ? __ldg(p.cw + c) : 1.f;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&done_bar[s], NWG); mbar_init(&empty_bar[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    constexpr int kBlocks = TP * (int)sizeof(T) / kCsBoxBytes;     // TMA boxes per tile
    const int pad_words = (CP - C) * (kCsBoxBytes / 4);            // per block: rows C..CP-1 (swizzling permutes inside a row)
    for (int s = 0; s < NS; ++s)
      for (int bx = 0; bx < kBlocks; ++bx)
        for (int i = tid; i < pad_words; i += blockDim.x)
          reinterpret_cast<uint32_t*>(smem_raw + (size_t)s * p.stage_bytes + (size_t)bx * p.block_bytes + (size_t)C * kCsBoxBytes)[i] =
              neg_inf_word<T>();
  }
  __syncthreads();
  const long long t0 = (long long)blockIdx.x * p.tiles_per_cta;
  const long long t1 = (t0 + p.tiles_per_cta < p.total_tiles) ? t0 + p.tiles_per_cta : p.total_tiles;

  if (warp == NW) {
    cs_producer<T, true, true>(p, &tmap, smem_raw, full_bar, empty_bar, t0, t1, TP, lane);
  } else if (warp == NW + 1) {
    // store warp: lane 0 hands the boxes of a finished tile to the TMA store engine; the stage goes back to the producer
    // as soon as the engine has READ it
    constexpr int BOXPX = CsCfg<T>::kBoxPx;
    CsWalker wk;
    wk.init(p, t1 - 1, 1, TP);
    const int nt = (int)(t1 - t0);
    StageRing ring;
    ring.init(0, NS, 1);
    for (int k = 0; k < nt; ++k, wk.backward(), ring.advance()) {
      const int s = ring.s;
      mbar_wait(&done_bar[s], ring.ph);
      if (lane == 0) {
        const CsTile tl = wk.tile();
        const int nbox = (tl.npx + BOXPX - 1) / BOXPX;
        unsigned char* stage = smem_raw + (size_t)s * p.stage_bytes;
        for (int bx = 0; bx < nbox; ++bx)
          tma_store_2d(&tmap_grad, (int)(tl.px0 + (long long)bx * BOXPX), tl.n * C, stage + (size_t)bx * p.block_bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(&empty_bar[s]);
      }
      __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    constexpr int PX = CsCfg<T>::PX, kPasses = CsCfg<T>::kPasses;
    const int j = lane & 7, g = lane >> 3;
    const int grp = warp / NWG;
    const int pxo = (warp % NWG) * kCsWarpPx;
    const bool want_ce = (p.flags & B200SEG_WANT_CE) != 0;
    const float god = p.dice_grad_out ? __ldg(p.dice_grad_out) : 1.f;
    float Gce = 0.f;
    if (want_ce) {
      Gce = p.ce_scale_host;
      if (p.ce_grad_out) Gce *= __ldg(p.ce_grad_out);
      if (p.ce_use_nvalid) {
        const double nv = (double)(long long)p.stats[B200SEG_ST_N_VALID];
        Gce = (float)((double)Gce / (nv + 1.1920928955078125e-07));
      }
    }
    float b[CPT];
    int n_cur = -1;
    CsWalker wk;
    wk.init(p, t1 - 1 - grp, G, TP);
    StageRing ring;
    ring.init(grp, NS, G);
    for (int k = grp; k < (int)(t1 - t0); k += G, wk.backward(), ring.advance()) {
      const int s = ring.s;
      const CsTile tl = wk.tile();
      if (tl.n != n_cur) {
        // new image: the group's 8 warps reload its coefficient table (two named barriers; once or twice per CTA)
        n_cur = tl.n;
        const int gt = tid - grp * (NWG * 32);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(NWG * 32) : "memory");
        for (int c = gt; c < C; c += NWG * 32) {
          const float2 ab = __ldg(reinterpret_cast<const float2*>(p.dice_coef) + (size_t)tl.n * C + c);
          coef_s[grp][0][c] = god * ab.x;
          coef_s[grp][1][c] = 2.f * god * ab.y;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(NWG * 32) : "memory");
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          const int c = i * kCsSlices + j;
          b[i] = c < C ? coef_s[grp][1][c] : 0.f;
        }
      }
      mbar_wait(&full_bar[s], ring.ph);
      unsigned char* stage = smem_raw + (size_t)s * p.stage_bytes;

      // per-pixel scalars of the pixel this lane owns (lane j < 4: pass j / PX, v = j % PX): issued first, consumed
      // after the class sweep
      float extra = 0.f, py = 0.f, da = 0.f, kk = 0.f, by = 0.f;
      int ycc = 0;
      const int t_px = pxo + (j / PX) * 4 * PX + g * PX + (j % PX);
      const bool owner = (j < 4) && t_px < tl.npx;
      if (owner) {
        const int yy = cs_label32(stage + p.label_off, p.label_dtype, t_px);
        const bool valid = (yy != p.ignore32) && (unsigned)yy < (unsigned)C;
        ycc = yy < 0 ? 0 : (yy >= C ? C - 1 : yy);
        const size_t gpx = (size_t)tl.n * p.HW + tl.px0 + t_px;
        if (want_ce && valid) kk = Gce * (p.pw ? __ldg(p.pw + gpx) : 1.f) * cw_s[ycc];
        if (yy != p.dice_ignore32) da = coef_s[grp][0][ycc];
        by = coef_s[grp][1][ycc];
        const float lse_own = reinterpret_cast<const float*>(stage + p.lse_off)[t_px];
        const float zy = to_float<T>(*reinterpret_cast<const T*>(stage + cs_offset<T>(p, ycc, t_px)));
        py = ex2(fmaf(zy, kLog2e, -lse_own * kLog2e));
        extra = -fmaf(da, py, kk);
      }
      float Down = 0.f;                                     // D of the owned pixel (set in its pass)
#pragma unroll
      for (int h = 0; h < kPasses; ++h) {
        const int pxh = pxo + h * 4 * PX + g * PX;
        unsigned char* col = stage + cs_offset<T>(p, j, pxh);
        float nl[PX], D[PX], Db[PX];
#pragma unroll
        for (int v = 0; v < PX; ++v) {
          nl[v] = -reinterpret_cast<const float*>(stage + p.lse_off)[pxh + v] * kLog2e;
          D[v] = 0.f;
          Db[v] = 0.f;
        }
        float pr[CPT][PX];
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          const CsRow<T> zr = cs_load<T>(col + (size_t)i * (kCsSlices * kCsBoxBytes));
          float z[PX];
          cs_unpack<T>(zr, z);
#pragma unroll
          for (int v = 0; v < PX; ++v) {
            pr[i][v] = ex2(fmaf(z[v], kLog2e, nl[v]));
            if (i & 1) Db[v] = fmaf(pr[i][v] * b[i], pr[i][v], Db[v]);   // two partial sums: shorter dependent chains
            else D[v] = fmaf(pr[i][v] * b[i], pr[i][v], D[v]);
          }
        }
#pragma unroll
        for (int v = 0; v < PX; ++v) D[v] += Db[v];
        // the owner adds -(da p_y + kk): after the tree every lane of the pixel holds D = dot - kk = -sub
#pragma unroll
        for (int v = 0; v < PX; ++v) D[v] += (j == h * PX + v) ? extra : 0.f;
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
          for (int v = 0; v < PX; ++v) D[v] += __shfl_xor_sync(0xffffffffu, D[v], o);
        }
#pragma unroll
        for (int v = 0; v < PX; ++v) Down = (j == h * PX + v) ? D[v] : Down;
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          float gq[PX];
#pragma unroll
          for (int v = 0; v < PX; ++v) gq[v] = pr[i][v] * fmaf(pr[i][v], b[i], -D[v]);
          if (i < CPT - 1 || i * kCsSlices + j < C) cs_store<T>(col + (size_t)i * (kCsSlices * kCsBoxBytes), gq);   // pad rows stay -inf
        }
      }
      __syncwarp();
      // one-hot term: the label's class is re-stored by the owner (after the slice lanes, same warp) with -(p_y da + kk)
      if (owner && (da != 0.f || kk != 0.f))
        *reinterpret_cast<T*>(stage + cs_offset<T>(p, ycc, t_px)) = from_float<T>(py * fmaf(py, by, -Down) - fmaf(py, da, kk));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&done_bar[s]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct CsGeom {
  int cpt, groups, stages, stage_bytes, block_bytes, label_off, lse_off;
  size_t smem_fwd, smem_bwd;
};

static int cs_cpt_for(int C) {
  if (C <= 64) return 8;
  if (C <= 96) return 12;
  if (C <= 128) return 16;
  if (C <= 152) return 19;
  return 0;
}

static bool cs_geometry(int C, int elem, CsGeom* g) {
  const int cpt = cs_cpt_for(C);
  if (!cpt) return false;
  constexpr int TP = kCsGroupWarps * kCsWarpPx;
  g->cpt = cpt;
  g->groups = elem == 2 ? 2 : 1;
  g->block_bytes = cpt * kCsSlices * kCsBoxBytes;                      // CP rows x 128 bytes: a multiple of 1024
  const int blocks = TP * elem / kCsBoxBytes;
  g->label_off = blocks * g->block_bytes;
  g->lse_off = g->label_off + TP * 8;
  g->stage_bytes = ((g->lse_off + TP * 4 + 1023) / 1024) * 1024;      // stages stay 1024-byte aligned (swizzle phase)
  const size_t bins = (size_t)kCsGroupWarps * g->groups * 2 * C * sizeof(float);
  const size_t budget = (size_t)(224 * 1024);
  int st = (int)((budget - bins) / g->stage_bytes);
  if (st > kCsMaxStages) st = kCsMaxStages;
  if (st < 2) return false;
  g->stages = st;
  g->smem_fwd = (size_t)st * g->stage_bytes + bins + 1024;   // + alignment slack
  g->smem_bwd = (size_t)st * g->stage_bytes + 1024;
  return true;
}

// 16-byte tileable CE + Dice (exponent 2) at label resolution with 32 < C <= 152
bool cs_supported(const void* logits, const void* labels, const void* lse, const void* grad, int logit_dtype, int label_dtype,
                  int C, long long HW, float dice_exponent, int dice_mode, bool want_dice) {
  if (C <= 32 || !cs_cpt_for(C) || HW < 1) return false;
  if (want_dice && (dice_mode != B200SEG_MODE_DICE || dice_exponent != 2.f)) return false;
  const int elem = logit_bytes(logit_dtype), lb = label_bytes(label_dtype);
  if (!aligned16(logits) || !aligned16(labels) || !aligned16(lse) || (grad && !aligned16(grad))) return false;
  if ((HW * elem) % 16 || (HW * lb) % 16 || (HW * 4) % 16) return false;
  CsGeom g;
  return cs_geometry(C, elem, &g);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn cs_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// (N*C, HW) view of an NCHW tensor; box = (128 bytes of pixels) x (C rows), 128B swizzle, out-of-range pixels read as 0
static int cs_make_map(CUtensorMap* map, const void* base, int logit_dtype, int N, int C, long long HW) {
  EncodeTiledFn enc = cs_encode_fn();
  B200SEG_REQUIRE(enc != nullptr, "class-sliced pipeline: cuTensorMapEncodeTiled is not available from this driver");
  const int elem = logit_bytes(logit_dtype);
  const CUtensorMapDataType dt = logit_dtype == B200SEG_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                               : (logit_dtype == B200SEG_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  const cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)N * (cuuint64_t)C};
  const cuuint64_t strides[1] = {(cuuint64_t)HW * (cuuint64_t)elem};
  const cuuint32_t box[2] = {(cuuint32_t)(kCsBoxBytes / elem), (cuuint32_t)C};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200SEG_REQUIRE(r == CUDA_SUCCESS, "class-sliced pipeline: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

template <typename T, int CPT, bool BWD> static int cs_launch_t(CsParams p, const CsGeom& g, int logit_dtype, cudaStream_t st) {
  constexpr int G = sizeof(T) == 2 ? 2 : 1;   // 16-bit: two consumer groups of 8 warps share one CTA
  constexpr int NW = kCsGroupWarps * G;
  long long grid = kSMs;
  if (grid > p.total_tiles) grid = p.total_tiles;
  p.tiles_per_cta = (p.total_tiles + grid - 1) / grid;
  grid = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  CUtensorMap map_in;
  if (int e = cs_make_map(&map_in, p.logits, logit_dtype, p.N, p.C, p.HW)) return e;
  if constexpr (BWD) {
    CUtensorMap map_out;
    if (int e = cs_make_map(&map_out, p.grad, logit_dtype, p.N, p.C, p.HW)) return e;
    auto k = cs_bwd_kernel<T, CPT, G>;
    if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(k), (int)g.smem_bwd)) return e;
    k<<<(unsigned)grid, (NW + 2) * 32, g.smem_bwd, st>>>(p, map_in, map_out);
    count_launch();
    return check_launch("cs_bwd_kernel");
  } else {
    auto k = cs_fwd_kernel<T, CPT, G>;
    if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(k), (int)g.smem_fwd)) return e;
    k<<<(unsigned)grid, (NW + 1) * 32, g.smem_fwd, st>>>(p, map_in);
    count_launch();
    return check_launch("cs_fwd_kernel");
  }
}

template <typename T, bool BWD> static int cs_launch(const CsParams& p, const CsGeom& g, int logit_dtype, cudaStream_t st) {
  switch (g.cpt) {
    case 8: return cs_launch_t<T, 8, BWD>(p, g, logit_dtype, st);
    case 12: return cs_launch_t<T, 12, BWD>(p, g, logit_dtype, st);
    case 16: return cs_launch_t<T, 16, BWD>(p, g, logit_dtype, st);
    case 19: return cs_launch_t<T, 19, BWD>(p, g, logit_dtype, st);
  }
  set_error("class-sliced pipeline: unsupported class count %d", p.C);
  return 1;
}

template <bool BWD> static int cs_dispatch(CsParams p, int logit_dtype, cudaStream_t st) {
  CsGeom g;
  const int elem = logit_bytes(logit_dtype);
  B200SEG_REQUIRE(cs_geometry(p.C, elem, &g), "class-sliced pipeline: unsupported shape (C=%d)", p.C);
  constexpr int TP = kCsGroupWarps * kCsWarpPx;
  p.stages = g.stages; p.stage_bytes = g.stage_bytes; p.block_bytes = g.block_bytes; p.label_off = g.label_off; p.lse_off = g.lse_off;
  p.tiles_per_image = (int)((p.HW + TP - 1) / TP);
  p.total_tiles = (long long)p.tiles_per_image * p.N;
  switch (logit_dtype) {
    case B200SEG_F32: return cs_launch<float, BWD>(p, g, logit_dtype, st);
    case B200SEG_BF16: return cs_launch<__nv_bfloat16, BWD>(p, g, logit_dtype, st);
    case B200SEG_F16: return cs_launch<__half, BWD>(p, g, logit_dtype, st);
  }
  set_error("unsupported logit dtype %d", logit_dtype);
  return 1;
}

static int cs_fit32(long long v) { return (v >= -2147483647LL && v <= 2147483647LL) ? (int)v : kCsNever; }

int cs_fwd_dispatch(const b200seg_loss_desc* d, cudaStream_t st) {
  CsParams p = {};
  p.logits = d->logits; p.labels = d->labels; p.pw = d->pixel_weight; p.cw = d->ce_class_weight;
  p.lse = (d->flags & B200SEG_WANT_LSE) ? d->lse : nullptr;
  p.stats = reinterpret_cast<unsigned long long*>(d->stats);
  p.dice_part = (d->flags & B200SEG_WANT_DICE) ? d->dice_part : nullptr;
  p.label_dtype = d->label_dtype; p.label_bytes = label_bytes(d->label_dtype);
  p.N = d->N; p.C = d->C; p.HW = (long long)d->H * d->W;
  p.flags = d->flags;
  p.ignore32 = cs_fit32(d->ignore_index); p.dice_ignore32 = cs_fit32(d->dice_ignore_index);
  p.acc_has_ignore = d->acc_has_ignore; p.acc_ignore32 = cs_fit32(d->acc_ignore_index);
  return cs_dispatch<false>(p, d->logit_dtype, st);
}

int cs_bwd_dispatch(const b200seg_loss_bwd_desc* d, cudaStream_t st) {
  CsParams p = {};
  p.logits = d->logits; p.labels = d->labels; p.pw = d->pixel_weight; p.cw = d->ce_class_weight;
  p.lse = const_cast<float*>(d->lse);
  p.stats = const_cast<unsigned long long*>(reinterpret_cast<const unsigned long long*>(d->stats));
  p.dice_coef = d->dice_coef; p.dice_grad_out = d->dice_grad_out; p.ce_grad_out = d->ce_grad_out;
  p.grad = d->grad_logits;
  p.label_dtype = d->label_dtype; p.label_bytes = label_bytes(d->label_dtype);
  p.N = d->N; p.C = d->C; p.HW = (long long)d->H * d->W;
  p.flags = d->flags;
  p.ignore32 = cs_fit32(d->ignore_index); p.dice_ignore32 = cs_fit32(d->dice_ignore_index);
  p.ce_scale_host = d->ce_scale_host; p.ce_use_nvalid = d->ce_use_nvalid;
  return cs_dispatch<true>(p, d->logit_dtype, st);
}

}  // namespace b200seg
