// Lovasz-Softmax / Lovasz hinge loss (the tail of SURVEY 8 f4), sm_100a. Every kernel in this file is hand-written;
// no library sort (round 1 called cub::DeviceRadixSort once per class: 73 % of the time of this row).
//
// Replaces models/losses/lovasz_loss.py:26-234. Per class c the reference materialises softmax(N,C,H,W), permutes it to
// (P,C), compacts the valid pixels (boolean index + nonzero), then for every class builds fg = (labels == c),
// errors = |fg - p_c|, torch.sort(errors, descending), gathers fg through the permutation, forms the Jaccard gradient
// with two fp32 cumsums and a shifted difference (lovasz_grad, :26-39) and takes a dot product: ~12 launches and ~10
// (P,)-sized temporaries per class in a Python loop, a host sync per class for 'present' (:153), and an autograd graph
// that walks all of it backwards.
//
// Here one SEGMENT is a (class c, image group g) pair — one group = the whole batch, or one image when per_image=True —
// and ALL segments of a batch of classes go through every kernel together (grid.y / a ticket = the segment), so that a
// launch covers tens of millions of items and streams at HBM rate instead of paying launch tails per class:
//   lovasz_prep_kernel   labels (any dtype) -> int16 class ids.
//   lovasz_keys_kernel   p_c = ex2(z_c*log2e - lse*log2e) from the logits row of class c and the per-pixel log-sum-exp of
//                        ONE forward pass (b200seg_loss_fwd, WANT_LSE); error and foreground bit packed into one 32-bit
//                        sort key:  key = ((bits(e) + 1) << 1) | fg  (e >= 0, so its fp32 bit pattern is monotone;
//                        lossless); ignored pixels get key 0 and sink to the end of the descending order. The same
//                        kernel counts the digit histograms of EVERY sort pass in shared memory (a CTA covers 32 K
//                        items, then adds its non-zero counts to the segment's histograms), so the keys are not read
//                        again for them.
//   lov_hist_scan_kernel exclusive scan of each (pass, segment) histogram -> first output slot of every digit value.
//   lov_sort_pass_kernel one least-significant-digit radix pass, stable, over (key, pixel index) pairs — keys only when
//                        no gradient is wanted — for all segments at once. Single read / single write per pass: a tile
//                        (256 threads x 16) ranks its items against warp-private counters (lanes with equal digits find
//                        each other through one ballot per digit bit),
//                        publishes its digit counts, orders the tile in shared memory while the counts of the tiles
//                        before it are collected by a decoupled look-back (tiles take their index from a ticket
//                        counter, so every tile waited for is already resident), and writes runs of equal digits.
//   lovasz_count_kernel / lovasz_tilescan_kernel / lovasz_grad_kernel
//                        exclusive scan of the foreground bits over the sorted order (tile counts -> one-CTA scan ->
//                        per-tile rescan), then with EXACT integer counts cum_i = #fg in [0,i], I_i = gts - cum_i,
//                        U_i = gts + (i + 1 - cum_i) the Jaccard increment in closed form
//                            g_i = 1/U_i              (fg_i = 1)
//                            g_i = I_i/(U_i (U_i-1))  (fg_i = 0),   g_0 = 1 - I_0/U_0
//                        (the reference's fp32 J_i - J_{i-1} cancels catastrophically: for P = 4 M its increments carry
//                        ~25 % noise; its cumsums stop being exact at 2^24), loss_c = sum e_i g_i, and
//                        dloss_c/dp_c = -+g_i scattered to the pixel's slot of G (C,N,HW) f32 (class-major: a segment
//                        is one contiguous slice, the scatter index is the sorted value itself).
//   lovasz_finalize_kernel  mean over the present / all / listed classes, class weights, per-image reduction
//                        (weight_reduce_loss), and the coefficient table of the backward. No host sync anywhere.
//   lovasz_bwd_kernel    softmax Jacobian: grad_z_j = up * p_j (a_j - sum_c a_c p_c), a_c = coef_c G_c.
// The binary hinge variant (lovasz_hinge_flat :69-91) runs the same pipeline on errors 1 - z*sign with an order-preserving
// float->uint key and the foreground bit in the value's top bit.
//
// Ties: the loss is invariant to the order inside a block of equal errors (the block's increments telescope); the
// gradient is not (neither is the reference's: torch.sort's order among ties is unspecified). The sort here is stable
// and deterministic: equal errors keep pixel order.
#include <cstdlib>

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

// ---------------------------------------------------------------------------------------------- radix sort geometry
// Descending order = ascending order of the complemented key: digit(pass) = (~key >> pass*RB) & (2^RB - 1).
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
template <int RB> struct SortGeo {
  static constexpr int NB = 1 << RB;                 // digit values per pass
  static constexpr int NP = (32 + RB - 1) / RB;      // passes over a 32-bit key
  static constexpr int BPT = NB / kSortThreads;      // digit values owned by a thread in the scans (RB >= 8)
  static_assert(RB >= 8 && RB <= 11, "digit width");
};
__device__ __forceinline__ uint32_t sort_digit(uint32_t key, int shift, uint32_t mask) { return ((~key) >> shift) & mask; }

// Lanes of the warp holding the same RB-bit digit: one ballot per digit bit, 4 instructions each (bit test, ballot, mask
// select, one LOP3: peers &= bit ? ballot : ~ballot). match.any does the same in one instruction but measured slower on
// B200 (its cost grows with the number of distinct values in the warp: 2.42 ms against 2.00 ms for the four passes).
template <int K> __device__ __forceinline__ unsigned peers_step(unsigned peers, uint32_t d) {
  unsigned r;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " .reg .b32 t, b;\n"
      " and.b32 t, %2, %3;\n"
      " setp.ne.u32 p, t, 0;\n"
      " vote.sync.ballot.b32 b, p, 0xffffffff;\n"
      " selp.b32 t, 0xffffffff, 0, p;\n"
      " lop3.b32 %0, %1, b, t, 0x90;\n"
      "}\n"
      : "=r"(r)
      : "r"(peers), "r"(d), "n"(1u << K));
  return r;
}
template <int RB> __device__ __forceinline__ unsigned digit_peers(uint32_t d) {
  unsigned peers = 0xffffffffu;
  peers = peers_step<0>(peers, d); peers = peers_step<1>(peers, d); peers = peers_step<2>(peers, d);
  peers = peers_step<3>(peers, d); peers = peers_step<4>(peers, d); peers = peers_step<5>(peers, d);
  peers = peers_step<6>(peers, d); peers = peers_step<7>(peers, d);
  if constexpr (RB > 8) peers = peers_step<8>(peers, d);
  if constexpr (RB > 9) peers = peers_step<9>(peers, d);
  if constexpr (RB > 10) peers = peers_step<10>(peers, d);
  return peers;
}

// ---------------------------------------------------------------------------------------------- sort keys + histograms
struct LovKeysParams {
  const void* logits;     // (N,C,HW) multi-class, (N,HW) binary
  const float* lse;       // (N,HW), multi-class only
  const int16_t* lab16;   // (N,HW)
  uint32_t* keys;         // (segments, len)
  uint32_t* vals;         // same count, or NULL (keys only)
  uint32_t* hist;         // (passes, segments, 2^RB), zeroed by the caller
  long long HW, len;      // pixels per image, items per segment
  int C, c0;              // classes of the tensor, first class of this batch of classes
  int groups, nseg;       // image groups per class (1, or N when per_image), segments = classes in the batch * groups
  int chunk;              // items per CTA (a multiple of 256 * V)
};

// grid (chunks per segment, segments). Segment s = (class c0 + s / groups, group s % groups); its items are the pixels
// g * len + i of the flat (N,HW) maps. value = index inside the segment (binary: | foreground << 31).
template <typename T, int V, bool BINARY, int RB>
__global__ void __launch_bounds__(kSortThreads) lovasz_keys_kernel(const LovKeysParams p) {
  using G = SortGeo<RB>;
  extern __shared__ __align__(16) unsigned char lov_smem[];
  uint32_t* hist_s = reinterpret_cast<uint32_t*>(lov_smem);             // [NP][NB] digit counts of this CTA's chunk
  const int tid = threadIdx.x;
  for (int i = tid; i < G::NP * G::NB; i += kSortThreads) hist_s[i] = 0u;
  __syncthreads();
  const int s = blockIdx.y;
  const int cj = s / p.groups, g = s - cj * p.groups;
  const int c = p.c0 + cj;
  const long long start = (long long)blockIdx.x * p.chunk;
  const long long end = min(p.len, start + (long long)p.chunk);
  const size_t seg0 = (size_t)s * (size_t)p.len;
  for (long long base = start; base < end; base += kSortThreads * V) {
    const long long i = base + (long long)tid * V;
    const bool active = i < end;                       // len % V == 0: a thread's V items are all inside or all outside
    uint32_t key[V], val[V];
    if (active) {
      const long long fp = (long long)g * p.len + i;   // flat pixel in (N,HW)
      const long long n = p.groups > 1 ? (long long)g : fp / p.HW;
      const long long hw = fp - n * p.HW;
      const T* zp = reinterpret_cast<const T*>(p.logits) + (BINARY ? (size_t)fp : ((size_t)n * p.C + c) * p.HW + hw);
      float z[V];
      load_vec<T, V>(zp, z);
      int lab[V];
      if constexpr (V == 4) {
        const uint2 r = ld_stream8(p.lab16 + fp);
        lab[0] = (int16_t)(r.x & 0xffffu); lab[1] = (int16_t)(r.x >> 16);
        lab[2] = (int16_t)(r.y & 0xffffu); lab[3] = (int16_t)(r.y >> 16);
      } else {
        lab[0] = p.lab16[fp];
      }
      if constexpr (BINARY) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const uint32_t fg = lab[v] != 0 && lab[v] != kLovIgnored;       // labels are 0 / 1 (:78-79)
          const float sign = fg ? 1.f : -1.f;
          key[v] = lab[v] == kLovIgnored ? 0u : hinge_key(1.f - z[v] * sign);
          val[v] = ((uint32_t)i + v) | (fg << 31);
        }
      } else {
        float l[V];
        load_vec<float, V>(p.lse + fp, l);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float pc = ex2(fmaf(z[v], kLog2e, -l[v] * kLog2e));
          const uint32_t fg = lab[v] == c;
          const float e = fabsf((fg ? 1.f : 0.f) - pc);
          key[v] = lab[v] == kLovIgnored ? 0u : (((__float_as_uint(e) + 1u) << 1) | fg);
          val[v] = (uint32_t)i + v;
        }
      }
      uint32_t* kout = p.keys + seg0 + i;
      if constexpr (V == 4) *reinterpret_cast<uint4*>(kout) = make_uint4(key[0], key[1], key[2], key[3]);
      else kout[0] = key[0];
      if (p.vals) {
        if constexpr (V == 4) *reinterpret_cast<uint4*>(p.vals + seg0 + i) = make_uint4(val[0], val[1], val[2], val[3]);
        else p.vals[seg0 + i] = val[0];
      }
#pragma unroll
      for (int ps = 0; ps < G::NP; ++ps)
#pragma unroll
        for (int v = 0; v < V; ++v) atomicAdd(hist_s + ps * G::NB + sort_digit(key[v], ps * RB, G::NB - 1), 1u);
    }
  }
  __syncthreads();
  for (int k = tid; k < G::NP * G::NB; k += kSortThreads) {
    const int ps = k / G::NB, b = k - ps * G::NB;
    const uint32_t t = hist_s[k];
    if (t) atomicAdd(p.hist + ((size_t)ps * p.nseg + s) * G::NB + b, t);
  }
}

// grid = passes * segments: exclusive scan of one histogram in place (first output slot of every digit value)
template <int RB>
__global__ void __launch_bounds__(kSortThreads) lov_hist_scan_kernel(uint32_t* __restrict__ hist) {
  using G = SortGeo<RB>;
  __shared__ uint32_t s_w[kSortWarps];
  uint32_t* h = hist + (size_t)blockIdx.x * G::NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t v[G::BPT], tsum = 0;
#pragma unroll
  for (int b = 0; b < G::BPT; ++b) { v[b] = h[tid * G::BPT + b]; tsum += v[b]; }
  uint32_t x = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_w[warp] = x;
  __syncthreads();
  uint32_t run = x - tsum;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) run += (w < warp) ? s_w[w] : 0u;
#pragma unroll
  for (int b = 0; b < G::BPT; ++b) { h[tid * G::BPT + b] = run; run += v[b]; }
}

// ---------------------------------------------------------------------------------------------- one radix pass
struct LovSortParams {
  const uint32_t* kin; uint32_t* kout;
  const uint32_t* vin; uint32_t* vout;     // PAIRS only
  const uint32_t* base;   // (segments, NB): first slot of every digit value inside the segment, this pass
  uint32_t* desc;         // (segments * tiles, NB): (count << 2) | state, zeroed by the caller; state 1 = this tile's
                          // own count, 2 = the count of this tile and all tiles before it in the segment
  uint32_t* ticket;       // zeroed by the caller
  long long len;          // items per segment (< 2^30)
  int tiles;              // tiles per segment
  int shift;
};

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int RB, int ITEMS, bool PAIRS> struct SortSmem {
  using G = SortGeo<RB>;
  static constexpr int TILE = kSortThreads * ITEMS;
  static constexpr int CNT_BYTES = kSortWarps * G::NB * 2;
  static constexpr int BYTES = CNT_BYTES + (PAIRS ? 2 : 1) * TILE * 4 + G::NB * 4;
};

// One tile = 256 threads x ITEMS consecutive items of one segment, order inside the tile = (warp, item, lane).
// FULL: every item of the tile exists (no bounds checks anywhere).
template <int RB, int ITEMS, bool PAIRS, bool FULL>
__device__ __forceinline__ void lov_sort_tile(const LovSortParams& p, unsigned char* lov_smem, uint32_t gt, int s, int tile,
                                              uint32_t* s_w) {
  using G = SortGeo<RB>;
  using SM = SortSmem<RB, ITEMS, PAIRS>;
  constexpr int NB = G::NB, BPT = G::BPT, TILE = SM::TILE;
  constexpr uint32_t MASK = NB - 1;
  static_assert(ITEMS % 2 == 0, "ranks are kept in 16-bit pairs");
  uint16_t* cnt = reinterpret_cast<uint16_t*>(lov_smem);                         // [warps][NB]
  uint32_t* exk = reinterpret_cast<uint32_t*>(lov_smem + SM::CNT_BYTES);         // [TILE] keys in tile order
  uint32_t* exv = exk + TILE;                                                    // [TILE] values (PAIRS)
  uint32_t* gdelta = exk + (PAIRS ? 2 : 1) * TILE;                               // [NB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t seg0 = (size_t)s * (size_t)p.len;
  const uint32_t tile0 = (uint32_t)tile * TILE;
  const uint32_t nvalid = FULL ? TILE : (uint32_t)(p.len - tile0);
  const int shift = p.shift;

  // ---- load (coalesced: lane-strided inside the warp's slice). Items past the end take key 0: the largest digit of the
  // pass, and being last in tile order they land behind every real item.
  const uint32_t w0 = (uint32_t)warp * (32 * ITEMS) + lane;                      // first item of the thread inside the tile
  const uint32_t* kin = p.kin + seg0 + tile0 + w0;
  uint32_t dg[ITEMS], val[ITEMS];                                                // digit | key bits above it, see below
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) dg[j] = (FULL || w0 + j * 32 < nvalid) ? __ldcs(kin + j * 32) : 0u;
  if constexpr (PAIRS) {
    const uint32_t* vin = p.vin + seg0 + tile0 + w0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) val[j] = (FULL || w0 + j * 32 < nvalid) ? __ldcs(vin + j * 32) : 0u;
  }
  // ---- rank: position of the item among the warp's items with the same digit (16-bit pairs)
  uint32_t rk[ITEMS / 2];
  uint16_t* wcnt = cnt + warp * NB;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t d = sort_digit(dg[j], shift, MASK);
    const unsigned peers = digit_peers<RB>(d);
    const unsigned lower = peers & lt;
    uint32_t old = 0;
    if (lower == 0u) {                                 // lowest lane of the group
      old = wcnt[d];
      wcnt[d] = (uint16_t)(old + __popc(peers));
    }
    old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1) + __popc(lower);
    if (j & 1) rk[j / 2] = __byte_perm(rk[j / 2], old, 0x5410);
    else rk[j / 2] = old;
    __syncwarp();
  }
  __syncthreads();

  // ---- counters -> tile totals per digit value -> first slot of every digit value in the tile; the warp counters are
  // replaced by the first slot of the warp's items of that digit value (16 bits: a tile has at most 2^16 items)
  static_assert(TILE <= 65536, "tile slots are kept in 16 bits");
  uint32_t tot[BPT];
  uint16_t cw[kSortWarps][BPT];
#pragma unroll
  for (int b = 0; b < BPT; ++b) tot[b] = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) {
#pragma unroll
    for (int b = 0; b < BPT; ++b) {
      cw[w][b] = cnt[w * NB + tid * BPT + b];
      tot[b] += cw[w][b];
    }
  }
  uint32_t tsum = 0;
#pragma unroll
  for (int b = 0; b < BPT; ++b) tsum += tot[b];
  uint32_t x = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_w[warp] = x;
  // publish this tile's counts before anything else (the tiles after it wait for them)
  const uint32_t dmax = (0xffffffffu >> shift) & MASK;
  const uint32_t ninv = TILE - nvalid;
  uint32_t* drow = p.desc + (size_t)gt * NB + tid * BPT;
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    const uint32_t t = tot[b] - ((!FULL && (uint32_t)(tid * BPT + b) == dmax) ? ninv : 0u);
    st_relaxed_u32(drow + b, (t << 2) | (tile == 0 ? 2u : 1u));
  }
  __syncthreads();
  uint32_t bstart[BPT];
  {
    uint32_t run = x - tsum;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) run += (w < warp) ? s_w[w] : 0u;
#pragma unroll
    for (int b = 0; b < BPT; ++b) {
      bstart[b] = run;
      uint32_t r = run;
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        cnt[w * NB + tid * BPT + b] = (uint16_t)r;
        r += cw[w][b];
      }
      run += tot[b];
      if (!FULL && (uint32_t)(tid * BPT + b) == dmax) tot[b] -= ninv;
    }
  }
  __syncthreads();

  // ---- order the tile in shared memory
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t d = sort_digit(dg[j], shift, MASK);
    const uint32_t q = wcnt[d] + ((j & 1) ? (rk[j / 2] >> 16) : (rk[j / 2] & 0xffffu));
    exk[q] = dg[j];
    if constexpr (PAIRS) exv[q] = val[j];
  }

  // ---- decoupled look-back: items with the same digit in the tiles before this one (of the same segment)
  uint32_t excl[BPT];
#pragma unroll
  for (int b = 0; b < BPT; ++b) excl[b] = 0;
  if (tile > 0) {
    uint32_t pending = (1u << BPT) - 1u;
    const uint32_t* row = drow;
    while (pending) {                                  // ends at tile 0 at the latest (it publishes state 2)
      row -= NB;
      uint32_t v[BPT];
      for (;;) {
        bool ready = true;
#pragma unroll
        for (int b = 0; b < BPT; ++b) {
          v[b] = ld_relaxed_u32(row + b);
          ready = ready && ((v[b] & 3u) != 0u || !((pending >> b) & 1u));
        }
        if (ready) break;
        __nanosleep(40);
      }
#pragma unroll
      for (int b = 0; b < BPT; ++b) {
        if ((pending >> b) & 1u) {
          excl[b] += v[b] >> 2;
          if ((v[b] & 3u) == 2u) pending &= ~(1u << b);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < BPT; ++b) st_relaxed_u32(drow + b, ((excl[b] + tot[b]) << 2) | 2u);
  }
  const uint32_t* brow = p.base + (size_t)s * NB + tid * BPT;
  const uint32_t seg0_32 = (uint32_t)seg0;             // all segments together hold < 2^31 items
#pragma unroll
  for (int b = 0; b < BPT; ++b) gdelta[tid * BPT + b] = seg0_32 + brow[b] + excl[b] - bstart[b];
  __syncthreads();

  // ---- write the runs: slot q of the tile goes to gdelta[its digit] + q
  uint32_t* kout = p.kout;
  uint32_t* vout = p.vout;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const uint32_t q = (uint32_t)tid + k * kSortThreads;
    if (FULL || q < nvalid) {
      const uint32_t kk = exk[q];
      const uint32_t a = gdelta[sort_digit(kk, shift, MASK)] + q;
      kout[a] = kk;
      if constexpr (PAIRS) vout[a] = exv[q];
    }
  }
}

template <int RB, int ITEMS, bool PAIRS>
__global__ void __launch_bounds__(kSortThreads, 4) lov_sort_pass_kernel(const LovSortParams p) {
  using SM = SortSmem<RB, ITEMS, PAIRS>;
  extern __shared__ __align__(16) unsigned char lov_smem[];
  __shared__ uint32_t s_ticket;
  __shared__ uint32_t s_w[kSortWarps];
  const int tid = threadIdx.x;
  if (tid == 0) s_ticket = atomicAdd(p.ticket, 1u);
  {
    uint4* z = reinterpret_cast<uint4*>(lov_smem);
    for (int i = tid; i < SM::CNT_BYTES / 16; i += kSortThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const uint32_t gt = s_ticket;                        // global tile number: tiles are taken in segment-major order
  const int s = (int)(gt / (uint32_t)p.tiles), tile = (int)(gt - (uint32_t)s * (uint32_t)p.tiles);
  if ((long long)(tile + 1) * SM::TILE <= p.len) lov_sort_tile<RB, ITEMS, PAIRS, true>(p, lov_smem, gt, s, tile, s_w);
  else lov_sort_tile<RB, ITEMS, PAIRS, false>(p, lov_smem, gt, s, tile, s_w);
}

// ---------------------------------------------------------------------------------------------- scan over the sorted order
// 8 consecutive sorted items of one segment
__device__ __forceinline__ void lov_load_tile(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                              long long i0, long long len, bool vec, uint32_t (&k)[kLovItems],
                                              uint32_t (&v)[kLovItems], bool want_vals) {
  if (vec && i0 + kLovItems <= len) {
    const uint4 a = *reinterpret_cast<const uint4*>(keys + i0), b = *reinterpret_cast<const uint4*>(keys + i0 + 4);
    k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
    if (want_vals) {
      const uint4 c = *reinterpret_cast<const uint4*>(vals + i0), d = *reinterpret_cast<const uint4*>(vals + i0 + 4);
      v[0] = c.x; v[1] = c.y; v[2] = c.z; v[3] = c.w; v[4] = d.x; v[5] = d.y; v[6] = d.z; v[7] = d.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kLovItems; ++j) {
      const bool in = i0 + j < len;
      k[j] = in ? keys[i0 + j] : 0u;
      v[j] = (in && want_vals) ? vals[i0 + j] : 0u;
    }
  }
}
template <bool BINARY> __device__ __forceinline__ uint32_t lov_fg(uint32_t key, uint32_t val) {
  if constexpr (BINARY) return key ? (val >> 31) : 0u;
  else return key & 1u;                                  // ignored items carry key 0
}

// grid (tiles per segment, segments): foreground count of every tile
template <bool BINARY>
__global__ void __launch_bounds__(kLovThreads) lovasz_count_kernel(const uint32_t* __restrict__ keys,
                                                                   const uint32_t* __restrict__ vals, long long len,
                                                                   int vec, uint32_t* __restrict__ tile_cnt) {
  __shared__ uint32_t s[kLovThreads / 32];
  const size_t seg0 = (size_t)blockIdx.y * len;
  const long long i0 = ((long long)blockIdx.x * kLovThreads + threadIdx.x) * kLovItems;
  uint32_t k[kLovItems], v[kLovItems];
  lov_load_tile(keys + seg0, vals ? vals + seg0 : nullptr, i0, len, vec != 0, k, v, BINARY);
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

// statistics slot of segment s = (class c0 + s / groups, group s % groups) inside seg_stats (n_groups, n_seg, 2); the
// pointer handed to the kernels already points at class c0 of group 0
__device__ __forceinline__ double* lov_seg_stat(double* base, int s, int groups, int n_seg) {
  const int cj = s / groups, g = s - cj * groups;
  return base + ((size_t)g * n_seg + cj) * 2;
}

// one CTA per segment: exclusive scan of its tile counts; writes the segment's foreground total (+1 = "segment processed")
__global__ void __launch_bounds__(1024) lovasz_tilescan_kernel(const uint32_t* __restrict__ tile_cnt_all,
                                                               uint32_t* __restrict__ tile_off_all, int nb,
                                                               double* __restrict__ seg_stat_base, int groups, int n_seg) {
  const uint32_t* tile_cnt = tile_cnt_all + (size_t)blockIdx.x * nb;      // one CTA per segment
  uint32_t* tile_off = tile_off_all + (size_t)blockIdx.x * nb;
  double* seg_stat = lov_seg_stat(seg_stat_base, blockIdx.x, groups, n_seg);
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
  const uint32_t* keys;
  const uint32_t* vals;       // NULL when no gradient is wanted (multi-class)
  const uint32_t* tile_off;   // (segments, tiles per segment)
  double* seg_stat;           // class c0, group 0: [0] loss accumulator, [1] gts + 1 (lov_seg_stat for the others)
  float* G;                   // NULL = forward only
  float* Gseg;                // slice of G where segment 0 starts: multi-class (C,N,HW) f32 at class c0, binary (N,HW)
  long long len;              // items per segment (consecutive segments are `len` apart in keys, vals and G)
  int groups, n_seg;
  int vec;
};

template <bool BINARY>
__global__ void __launch_bounds__(kLovThreads) lovasz_grad_kernel(const LovGradParams p) {
  __shared__ uint32_t s_w[kLovThreads / 32];
  __shared__ float s_l[kLovThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long i0 = ((long long)blockIdx.x * kLovThreads + threadIdx.x) * kLovItems;
  const bool want_vals = BINARY || p.G != nullptr;
  uint32_t k[kLovItems], v[kLovItems];
  const size_t seg0 = (size_t)blockIdx.y * p.len;
  double* seg_stat = lov_seg_stat(p.seg_stat, blockIdx.y, p.groups, p.n_seg);
  lov_load_tile(p.keys + seg0, p.vals ? p.vals + seg0 : nullptr, i0, p.len, p.vec != 0, k, v, want_vals);
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

constexpr int kLovRB = 8;                          // digit width: 4 passes (11 bits / 3 passes measured slower: 3.63 against 2.62 ms)
constexpr int kItemsRB8 = 16, kItemsRB11 = 32;     // items per thread of a sort tile
constexpr int kKeysChunk = 256 * 4 * 32;           // items per CTA of the keys kernel (4096 per warp: 16-bit counters)
constexpr long long kLovBudget = 4ll << 30;        // workspace the query asks for at most (classes go in batches beyond it)

static inline int lov_sort_tile(int rb) { return kSortThreads * (rb == 11 ? kItemsRB11 : kItemsRB8); }
static inline int lov_passes(int rb) { return (32 + rb - 1) / rb; }

struct LovWorkspace {
  uint32_t *keys_a, *keys_b, *vals_a, *vals_b, *tile_cnt, *tile_off;
  uint32_t *ticket, *hist, *desc;     // one zeroed region: tickets (one per pass), histograms, look-back descriptors
  void* zero_base;
  size_t zero_bytes, desc_pass_words, total;
  int tiles;
};

// S segments of len items each
static void lov_carve(int rb, long long len, int S, bool pairs, void* base, LovWorkspace* w) {
  const int nbins = 1 << rb, np = lov_passes(rb);
  const size_t items = (size_t)len * S;
  const size_t arr = align256(items * 4);
  const size_t nb = (size_t)((len + kLovTile - 1) / kLovTile) * S;
  const size_t tb = align256((nb + 1) * sizeof(uint32_t));
  w->tiles = (int)((len + lov_sort_tile(rb) - 1) / lov_sort_tile(rb));
  w->desc_pass_words = (size_t)S * w->tiles * nbins;
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  w->keys_a = reinterpret_cast<uint32_t*>(p + off); off += arr;
  w->keys_b = reinterpret_cast<uint32_t*>(p + off); off += arr;
  w->vals_a = reinterpret_cast<uint32_t*>(p + off); off += pairs ? arr : 0;
  w->vals_b = reinterpret_cast<uint32_t*>(p + off); off += pairs ? arr : 0;
  w->tile_cnt = reinterpret_cast<uint32_t*>(p + off); off += tb;
  w->tile_off = reinterpret_cast<uint32_t*>(p + off); off += tb;
  w->zero_base = p + off;
  const size_t z0 = off;
  w->ticket = reinterpret_cast<uint32_t*>(p + off); off += 256;
  w->hist = reinterpret_cast<uint32_t*>(p + off); off += align256((size_t)np * S * nbins * 4);
  w->desc = reinterpret_cast<uint32_t*>(p + off); off += align256((size_t)np * w->desc_pass_words * 4);
  w->zero_bytes = off - z0;
  w->total = off;
}
static size_t lov_total(long long len, int S, bool pairs) {
  LovWorkspace a;
  lov_carve(kLovRB, len, S, pairs, nullptr, &a);
  return a.total;
}
// largest batch of classes (1 .. n) whose workspace fits `bytes`; 0 if not even one class fits
static int lov_class_batch(long long len, int groups, int n, bool pairs, long long bytes) {
  int hi = n;
  if ((long long)hi * groups > 65535) hi = 65535 / groups;   // segments ride on grid.y
  if ((long long)hi * groups * len >= (1ll << 31)) hi = (int)(((1ll << 31) - 1) / (groups * len));   // 32-bit item index
  if (hi < 1) hi = 1;
  if ((long long)lov_total(len, groups, pairs) > bytes) return 0;
  int lo = 1;                                                // invariant: lo fits
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if ((long long)lov_total(len, mid * groups, pairs) <= bytes) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

long long lovasz_workspace_bytes(int N, int C, long long HW, int per_image, int pairs) {
  if (N <= 0 || HW <= 0 || C <= 0) return 256;
  const int groups = per_image ? N : 1;
  const long long len = per_image ? HW : (long long)N * HW;
  int J = lov_class_batch(len, groups, C, pairs != 0, kLovBudget);
  if (J < 1) J = 1;
  return (long long)lov_total(len, J * groups, pairs != 0);
}

template <int RB, int ITEMS, bool PAIRS>
static int lov_launch_pass(const LovSortParams& sp, unsigned grid, cudaStream_t st) {
  auto k = lov_sort_pass_kernel<RB, ITEMS, PAIRS>;
  constexpr int smem = SortSmem<RB, ITEMS, PAIRS>::BYTES;
  if (smem > 48 * 1024)
    if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(k), smem)) return e;
  k<<<grid, kSortThreads, smem, st>>>(sp);
  return 0;
}

template <typename T, int V, bool BINARY, int RB>
static int lov_launch_keys(const LovKeysParams& kp, dim3 grid, cudaStream_t st) {
  auto k = lovasz_keys_kernel<T, V, BINARY, RB>;
  constexpr int smem = SortGeo<RB>::NP * SortGeo<RB>::NB * 4;
  if (smem > 48 * 1024)
    if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(k), smem)) return e;
  k<<<grid, kSortThreads, smem, st>>>(kp);
  return 0;
}

// one batch of classes [c0, c0 + J): every kernel covers its J * groups segments
template <typename T, int RB>
static int lov_fwd_batch(const b200seg_lovasz_desc* d, int c0, int J, cudaStream_t st) {
  using G = SortGeo<RB>;
  constexpr int ITEMS = RB == 11 ? kItemsRB11 : kItemsRB8;
  const bool binary = d->binary != 0;
  const int n_seg = binary ? 1 : d->C;
  const long long HW = d->HW;
  const int groups = d->per_image ? d->N : 1;
  const long long len = d->per_image ? HW : (long long)d->N * HW;
  const int S = J * groups;
  const bool pairs = binary || d->G != nullptr;
  LovWorkspace w;
  lov_carve(RB, len, S, pairs, d->workspace, &w);
  B200SEG_REQUIRE((long long)w.total <= d->workspace_bytes, "lovasz_fwd: workspace of %lld bytes needed, %lld given",
                  (long long)w.total, (long long)d->workspace_bytes);
  B200SEG_CUDA(cudaMemsetAsync(w.zero_base, 0, w.zero_bytes, st));

  LovKeysParams kp;
  kp.logits = d->logits; kp.lse = d->lse; kp.lab16 = d->lab16;
  kp.keys = w.keys_a; kp.vals = pairs ? w.vals_a : nullptr; kp.hist = w.hist;
  kp.HW = HW; kp.len = len; kp.C = d->C; kp.c0 = c0; kp.groups = groups; kp.nseg = S; kp.chunk = kKeysChunk;
  const bool vec = HW % 4 == 0 && aligned16(d->logits) && aligned16(d->lab16) && (binary || aligned16(d->lse));
  dim3 kgrid((unsigned)((len + kKeysChunk - 1) / kKeysChunk), S);
  int e;
  if (vec) e = binary ? lov_launch_keys<T, 4, true, RB>(kp, kgrid, st) : lov_launch_keys<T, 4, false, RB>(kp, kgrid, st);
  else e = binary ? lov_launch_keys<T, 1, true, RB>(kp, kgrid, st) : lov_launch_keys<T, 1, false, RB>(kp, kgrid, st);
  if (e) return e;
  if ((e = check_launch("lovasz_keys_kernel"))) return e;
  lov_hist_scan_kernel<RB><<<G::NP * S, kSortThreads, 0, st>>>(w.hist);
  if ((e = check_launch("lov_hist_scan_kernel"))) return e;

  uint32_t *kin = w.keys_a, *kout = w.keys_b, *vin = w.vals_a, *vout = w.vals_b;
  for (int ps = 0; ps < G::NP; ++ps) {
    LovSortParams sp;
    sp.kin = kin; sp.kout = kout; sp.vin = pairs ? vin : nullptr; sp.vout = pairs ? vout : nullptr;
    sp.base = w.hist + (size_t)ps * S * G::NB;
    sp.desc = w.desc + (size_t)ps * w.desc_pass_words;
    sp.ticket = w.ticket + ps;
    sp.len = len; sp.tiles = w.tiles; sp.shift = ps * RB;
    const unsigned grid = (unsigned)((size_t)S * w.tiles);
    e = pairs ? lov_launch_pass<RB, ITEMS, true>(sp, grid, st) : lov_launch_pass<RB, ITEMS, false>(sp, grid, st);
    if (e) return e;
    if ((e = check_launch("lov_sort_pass_kernel"))) return e;
    uint32_t* t = kin; kin = kout; kout = t;
    t = vin; vin = vout; vout = t;
  }
  const uint32_t* ksorted = kin;
  const uint32_t* vsorted = pairs ? vin : nullptr;

  const int scan_vec = (S == 1 || len % 4 == 0) ? 1 : 0;      // 16-byte loads of the sorted arrays stay aligned
  const int nb = (int)((len + kLovTile - 1) / kLovTile);
  double* seg = d->seg_stats + (size_t)c0 * 2;
  dim3 sgrid(nb, S);
  if (binary) lovasz_count_kernel<true><<<sgrid, kLovThreads, 0, st>>>(ksorted, vsorted, len, scan_vec, w.tile_cnt);
  else lovasz_count_kernel<false><<<sgrid, kLovThreads, 0, st>>>(ksorted, nullptr, len, scan_vec, w.tile_cnt);
  lovasz_tilescan_kernel<<<S, 1024, 0, st>>>(w.tile_cnt, w.tile_off, nb, seg, groups, n_seg);
  LovGradParams gp;
  gp.keys = ksorted; gp.vals = vsorted; gp.tile_off = w.tile_off; gp.seg_stat = seg; gp.groups = groups; gp.n_seg = n_seg;
  gp.G = d->G; gp.len = len; gp.vec = scan_vec;
  gp.Gseg = d->G ? d->G + (size_t)(binary ? 0 : c0) * d->N * HW : nullptr;
  if (binary) lovasz_grad_kernel<true><<<sgrid, kLovThreads, 0, st>>>(gp);
  else lovasz_grad_kernel<false><<<sgrid, kLovThreads, 0, st>>>(gp);
  count_launch(5 + G::NP);
  return check_launch("lovasz scan kernels");
}

template <typename T, int RB>
static int lov_fwd_typed(const b200seg_lovasz_desc* d, cudaStream_t st) {
  const bool binary = d->binary != 0;
  const int groups = d->per_image ? d->N : 1;
  const long long len = d->per_image ? d->HW : (long long)d->N * d->HW;
  const bool pairs = binary || d->G != nullptr;
  B200SEG_REQUIRE(len < (1ll << 30), "lovasz_fwd: %lld items per segment exceed 2^30-1", len);
  const long long npx = (long long)d->N * d->HW;
  lovasz_prep_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(d->labels, d->label_dtype, npx, binary ? 2 : d->C,
                                                                    d->has_ignore, d->ignore_index, d->lab16);
  count_launch();
  if (int e = check_launch("lovasz_prep_kernel")) return e;
  if (binary) return lov_fwd_batch<T, RB>(d, 0, 1, st);
  if (d->classes_host) {                                  // explicit class list: one class per batch
    B200SEG_REQUIRE(lov_class_batch(len, groups, 1, pairs, d->workspace_bytes) >= 1,
                    "lovasz_fwd: workspace of %lld bytes is too small for one class", (long long)d->workspace_bytes);
    for (int j = 0; j < d->n_classes; ++j)
      if (int e = lov_fwd_batch<T, RB>(d, d->classes_host[j], 1, st)) return e;
    return 0;
  }
  const int J = lov_class_batch(len, groups, d->C, pairs, d->workspace_bytes);
  B200SEG_REQUIRE(J >= 1, "lovasz_fwd: workspace of %lld bytes is too small for one class", (long long)d->workspace_bytes);
  const int nbatch = (d->C + J - 1) / J;
  const int Jb = (d->C + nbatch - 1) / nbatch;            // even batches
  for (int c0 = 0; c0 < d->C; c0 += Jb)
    if (int e = lov_fwd_batch<T, RB>(d, c0, min(Jb, d->C - c0), st)) return e;
  return 0;
}

template <typename T> static int lov_fwd_rb(const b200seg_lovasz_desc* d, cudaStream_t st) {
  return lov_fwd_typed<T, kLovRB>(d, st);
}

int lovasz_fwd_dispatch(const b200seg_lovasz_desc* d, cudaStream_t st) {
  const int n_seg = d->binary ? 1 : d->C;
  const int n_groups = d->per_image ? d->N : 1;
  B200SEG_CUDA(cudaMemsetAsync(d->seg_stats, 0, (size_t)n_groups * n_seg * 2 * sizeof(double), st));
  if (d->N > 0 && d->HW > 0) {
    int e;
    switch (d->logit_dtype) {
      case B200SEG_F32: e = lov_fwd_rb<float>(d, st); break;
      case B200SEG_BF16: e = lov_fwd_rb<__nv_bfloat16>(d, st); break;
      default: e = lov_fwd_rb<__half>(d, st); break;
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
