// Single-pass soft-max cross-entropy forward AND backward at label resolution as a bulk-copy (TMA) pipeline, sm_100a.
//
// Replaces F.cross_entropy (models/losses/cross_entropy_loss.py:56-61), weight_reduce_loss (models/losses/utils.py:48-80),
// the top-1 accuracy (models/losses/accuracy.py:41-60) and their autograd backward — at least six passes over the
// (N,C,H,W) logits in ATen — with ONE read and ONE write of them (the algorithmic minimum for a gradient of the same
// shape), for class counts whose tile leaves room for two CTAs x three stages per SM (fp32: C <= 34, 16-bit: C <= 69).
//
// Structure. A persistent CTA walks tiles of 256 pixels x C classes. A producer warp issues one `cp.async.bulk`
// (global -> shared, completion on an mbarrier) per class row of the tile plus one for the label row, 3-4 stages ahead;
// four consumer warps (thread = 2 adjacent pixels) soft-max the tile out of shared memory, overwrite it in place with the gradient
// and hand it to the bulk-store engine (`cp.async.bulk` shared -> global, one per class row). No register ever holds
// data in flight, so the bytes in flight per SM are set by the stage count (up to ~200 KB), not by occupancy: the
// register-tile kernel this replaces (loss_rt.cuh, still used for shapes that are not 16-byte tileable) stalled at 73 %
// of the copy roofline because a thread had to hold its own loads.
//
// Bound: HBM. Algorithmic bytes per launch: 2*N*C*H*W*s + N*H*W*L.
#include "bulk_pipe.cuh"
#include "common.cuh"

namespace b200seg {

struct BulkParams {
  const void* logits;
  const void* labels;
  const float* cw;
  const float* ce_grad_out;
  unsigned long long* stats;
  void* grad;
  float ce_scale_host;
  int label_dtype, label_bytes;
  int N, C;
  long long HW;
  int tiles_per_image;
  long long total_tiles;
  int stages;
  int stage_bytes;       // C * kBulkPx * sizeof(T) + label row (kBulkPx * 8 bytes)
  long long ignore_index;
  int acc_has_ignore;
  long long acc_ignore;
  int want_acc;
  // forward-only mode (GRAD = false): optional per-pixel outputs, no gradient
  float* lse;            // (N,H,W) log-sum-exp or NULL
  float* loss_px;        // (N,H,W) loss_weight * per-pixel loss or NULL
  float lw;
};

constexpr int kBulkV = 2;                              // pixels per consumer thread (one 8-byte / 4-byte shared-memory access)
constexpr int kBulkConsumers = 128;                    // consumer threads
constexpr int kBulkPx = kBulkConsumers * kBulkV;       // pixels per tile
constexpr int kBulkThreads = kBulkConsumers + 64;      // + one load-producer warp + one store warp
constexpr int kBulkMaxStages = 4;

// GRAD = false is the forward-only form (validation loss, `reduction='none'`, the first pass of the two-pass plans): the
// same load pipeline, two sweeps over the tile instead of three, nothing written back — the store warp only hands the
// stage back to the producer; per-pixel log-sum-exp / loss maps are stored by the consumers (8 bytes per thread, coalesced).
template <typename T, bool GRAD>
__global__ void __launch_bounds__(kBulkThreads) ce_bulk_kernel(const BulkParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[kBulkMaxStages], done_bar[kBulkMaxStages], empty_bar[kBulkMaxStages];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, NS = p.stages;
  const long long HW = p.HW;
  const size_t row_stride = (size_t)kBulkPx * sizeof(T);            // one class row of a tile in shared memory
  const size_t label_off = (size_t)C * row_stride;                   // label row follows the class rows
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&done_bar[s], kBulkConsumers); mbar_init(&empty_bar[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  float loss_acc = 0.f;
  int n_valid = 0, n_correct = 0, n_bad = 0, n_acc = 0;

  if (warp == kBulkConsumers / 32) {
    // ===================== producer warp: one bulk copy per class row + one for the labels, NS tiles ahead
    int k = 0;
    StageRing ring;
    ring.init(0, NS, 1);
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++k, ring.advance()) {
      const int s = ring.s;
      if (k >= NS) mbar_wait(&empty_bar[s], ring.ph ^ 1);
      const int n = (int)(tile / p.tiles_per_image);
      const long long px0 = (tile - (long long)n * p.tiles_per_image) * kBulkPx;
      const int npx = (int)((HW - px0 < kBulkPx) ? HW - px0 : kBulkPx);
      const unsigned row_bytes = (unsigned)(npx * sizeof(T));
      const unsigned lab_bytes = (unsigned)(npx * p.label_bytes);
      unsigned char* stage = smem_raw + (size_t)s * p.stage_bytes;
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], (unsigned)C * row_bytes + lab_bytes);
      __syncwarp();
      const char* src = reinterpret_cast<const char*>(p.logits) + ((size_t)n * C * HW + px0) * sizeof(T);
      for (int c = lane; c < C; c += 32)
        bulk_g2s(stage + (size_t)c * row_stride, src + (size_t)c * HW * sizeof(T), row_bytes, &full_bar[s]);
      if (lane == 0)
        bulk_g2s(stage + label_off, reinterpret_cast<const char*>(p.labels) + ((size_t)n * HW + px0) * p.label_bytes, lab_bytes,
                 &full_bar[s]);
    }
  } else if (warp == kBulkConsumers / 32 + 1) {
    // ===================== store warp: one bulk store per class row once all consumers are done with the tile; a stage
    // goes back to the producer when the stores of the tile BEFORE have finished reading it (one tile of slack)
    int k = 0, prev_s = 0;
    StageRing ring;
    ring.init(0, NS, 1);
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++k, prev_s = ring.s, ring.advance()) {
      const int s = ring.s;
      mbar_wait(&done_bar[s], ring.ph);
      const int n = (int)(tile / p.tiles_per_image);
      const long long px0 = (tile - (long long)n * p.tiles_per_image) * kBulkPx;
      const int npx = (int)((HW - px0 < kBulkPx) ? HW - px0 : kBulkPx);
      const unsigned row_bytes = (unsigned)(npx * sizeof(T));
      unsigned char* stage = smem_raw + (size_t)s * p.stage_bytes;
      char* dst = reinterpret_cast<char*>(p.grad) + ((size_t)n * C * HW + px0) * sizeof(T);
      if (lane == 0) {
        if constexpr (GRAD) {
          for (int c = 0; c < C; ++c) bulk_s2g(dst + (size_t)c * HW * sizeof(T), stage + (size_t)c * row_stride, row_bytes);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (k >= 1) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            mbar_arrive(&empty_bar[prev_s]);
          }
        } else {
          (void)dst; (void)stage; (void)row_bytes; (void)prev_s;
          mbar_arrive(&empty_bar[s]);                       // every consumer has read the tile: the stage is free
        }
      }
      __syncwarp();
    }
    if (GRAD && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // ===================== consumer warps: thread = kBulkV adjacent pixels of the tile
    constexpr int V = kBulkV;
    const float Gs = p.ce_scale_host * (p.ce_grad_out ? __ldg(p.ce_grad_out) : 1.f);
    struct __align__(sizeof(T) * V) Pack { T v[V]; };
    int k = 0;
    StageRing ring;
    ring.init(0, NS, 1);
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++k, ring.advance()) {
      const int s = ring.s;
      mbar_wait(&full_bar[s], ring.ph);
      const int n = (int)(tile / p.tiles_per_image);
      const long long px0 = (tile - (long long)n * p.tiles_per_image) * kBulkPx;
      const int npx = (int)((HW - px0 < kBulkPx) ? HW - px0 : kBulkPx);
      unsigned char* stage = smem_raw + (size_t)s * p.stage_bytes;
      Pack* col = reinterpret_cast<Pack*>(stage) + tid;              // class c of this thread's pixels: col[c * 128]
      const int t0 = tid * V;
      if (t0 < npx) {     // npx is a multiple of 16 bytes / sizeof(T) >= V: a thread's pixels are all in or all out
        float m[V], zy[V], kk[V], nm[V];
        int idx[V], yc[V];
        bool valid[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const long long yy = smem_label(stage + label_off, p.label_dtype, t0 + v);
          const bool ign = (yy == p.ignore_index);
          const bool inr = (yy >= 0 && yy < (long long)C);
          valid[v] = !ign && inr;
          yc[v] = valid[v] ? (int)yy : 0;
          zy[v] = to_float<T>(reinterpret_cast<const T*>(stage)[(size_t)yc[v] * kBulkPx + t0 + v]);
          kk[v] = valid[v] ? (p.cw ? __ldg(p.cw + yc[v]) : 1.f) : 0.f;
          n_bad += (!ign && !inr);
          n_valid += !ign;
          const bool av = p.acc_has_ignore ? (yy != p.acc_ignore) : true;
          n_acc += av;
          idx[v] = (av && inr) ? (int)yy : -1;   // the "correct" test below: arg-max == label
        }
        // pass 1: max and its lowest index
        {
          const Pack z0 = col[0];
          int am[V];
#pragma unroll
          for (int v = 0; v < V; ++v) { m[v] = to_float<T>(z0.v[v]); am[v] = 0; }
#pragma unroll 4
          for (int c = 1; c < C; ++c) {
            const Pack z = col[c * kBulkConsumers];
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float zf = to_float<T>(z.v[v]);
              if (zf > m[v]) { m[v] = zf; am[v] = c; }
            }
          }
#pragma unroll
          for (int v = 0; v < V; ++v) { n_correct += (am[v] == idx[v]); nm[v] = -m[v] * kLog2e; }
        }
        // pass 2: sum of exponentials (fp32 tiles keep the exponentials in place)
        float ssum[V];
#pragma unroll
        for (int v = 0; v < V; ++v) ssum[v] = 0.f;
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
          Pack z = col[c * kBulkConsumers];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float e = ex2(fmaf(to_float<T>(z.v[v]), kLog2e, nm[v]));
            ssum[v] += e;
            if constexpr (GRAD && sizeof(T) == 4) z.v[v] = from_float<T>(e);
          }
          if constexpr (GRAD && sizeof(T) == 4) col[c * kBulkConsumers] = z;
        }
        float kg[V], rr[V], lsev[V], lpx[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float lse = m[v] + fast_log(ssum[v]);
          const float l = kk[v] * (lse - zy[v]);
          loss_acc += l;
          lsev[v] = lse;
          lpx[v] = l * p.lw;
          kg[v] = kk[v] * Gs;
          rr[v] = kg[v] * fast_rcp(ssum[v]);
        }
        if constexpr (!GRAD) {
          static_assert(V == 2, "one 8-byte store per thread");
          const size_t o = (size_t)n * HW + px0 + t0;
          if (p.lse) *reinterpret_cast<float2*>(p.lse + o) = make_float2(lsev[0], lsev[1]);
          if (p.loss_px) *reinterpret_cast<float2*>(p.loss_px + o) = make_float2(lpx[0], lpx[1]);
        }
        // pass 3: gradient k * (p - onehot) in place
        if constexpr (GRAD) {
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
          Pack z = col[c * kBulkConsumers];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            float e;
            if constexpr (sizeof(T) == 4) e = to_float<T>(z.v[v]);
            else e = ex2(fmaf(to_float<T>(z.v[v]), kLog2e, nm[v]));
            z.v[v] = from_float<T>(rr[v] * e);
          }
          col[c * kBulkConsumers] = z;
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
          if (valid[v])
            reinterpret_cast<T*>(stage)[(size_t)yc[v] * kBulkPx + t0 + v] = from_float<T>(rr[v] * ex2(fmaf(zy[v], kLog2e, nm[v])) - kg[v]);
        }
        }
      }
      // hand the tile to the store warp: generic-proxy writes -> async proxy, then arrive (no CTA-wide barrier: a consumer
      // warp goes straight on to the next tile)
      if constexpr (GRAD) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&done_bar[s]);
    }
  }
  cta_flush_stats(loss_acc, n_valid, n_correct, n_bad, n_acc, p.stats);
}

// ------------------------------------------------------------------------------------------------ host side
static size_t bulk_stage_bytes(int C, int elem) { return (size_t)C * kBulkPx * elem + (size_t)kBulkPx * 8; }

// 16-byte tileable: every class row of every tile starts on a 16-byte boundary and has a 16-byte multiple of bytes
bool bulk_supported(const void* logits, const void* labels, const void* grad, int logit_dtype, int label_dtype, int C,
                    long long HW, bool has_pixel_weight) {
  const int elem = logit_bytes(logit_dtype), lb = label_bytes(label_dtype);
  if (has_pixel_weight || C < 1 || HW < 1) return false;
  if (!aligned16(logits) || !aligned16(labels) || !aligned16(grad)) return false;
  if ((HW * elem) % 16 || (HW * lb) % 16) return false;
  // two resident CTAs x three stages must fit: with fewer consumer warps per SM the kernel turns consumer bound (measured:
  // bf16 C = 150 with 128-pixel tiles and one CTA per SM ran at 0.21 of the roofline, the two-pass kernels at 0.56)
  return 6 * bulk_stage_bytes(C, elem) <= 220 * 1024;   // fp32: C <= 34, 16-bit: C <= 69
}

template <typename T, bool GRAD = true> static int bulk_launch(BulkParams p, cudaStream_t st) {
  const size_t stage = bulk_stage_bytes(p.C, (int)sizeof(T));
  int stages = (int)((200 * 1024) / stage);
  if (stages > kBulkMaxStages) stages = kBulkMaxStages;
  // forward only: half the bytes per tile for the same arithmetic — the consumers are the limit, so the shared memory
  // goes into more resident CTAs (consumer warps) with two stages each rather than into deeper rings
  if (!GRAD && stages > 2) stages = 2;
  p.stages = stages;
  p.stage_bytes = (int)stage;
  const size_t smem = stage * stages;
  auto k = ce_bulk_kernel<T, GRAD>;
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(k), 200 * 1024)) return e;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 5) per_sm = 5;
  long long grid = (long long)kSMs * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  k<<<(unsigned)grid, kBulkThreads, smem, st>>>(p);
  count_launch();
  return check_launch("ce_bulk_kernel");
}

int bulk_fused_dispatch(const b200seg_loss_fused_desc* d, cudaStream_t st) {
  const b200seg_loss_desc* f = &d->fwd;
  BulkParams p = {};
  p.logits = f->logits; p.labels = f->labels; p.cw = f->ce_class_weight; p.ce_grad_out = d->grad_out;
  p.stats = reinterpret_cast<unsigned long long*>(f->stats); p.grad = d->grad_logits;
  p.ce_scale_host = d->grad_scale_host;
  p.label_dtype = f->label_dtype; p.label_bytes = label_bytes(f->label_dtype);
  p.N = f->N; p.C = f->C; p.HW = (long long)f->H * f->W;
  p.tiles_per_image = (int)((p.HW + kBulkPx - 1) / kBulkPx);
  p.total_tiles = (long long)p.tiles_per_image * p.N;
  p.ignore_index = f->ignore_index; p.acc_has_ignore = f->acc_has_ignore; p.acc_ignore = f->acc_ignore_index;
  switch (f->logit_dtype) {
    case B200SEG_F32: return bulk_launch<float>(p, st);
    case B200SEG_BF16: return bulk_launch<__nv_bfloat16>(p, st);
    case B200SEG_F16: return bulk_launch<__half>(p, st);
  }
  set_error("loss_fused: unsupported logit dtype %d", f->logit_dtype);
  return 1;
}

// Forward only (b200seg_loss_fwd without Dice, at label resolution, no pixel weights, 16-byte tileable): same statistics
// as ce_fwd_kernel, optional per-pixel log-sum-exp and loss maps.
bool bulk_fwd_supported(const b200seg_loss_desc* d) {
  if ((d->flags & B200SEG_WANT_DICE) || d->h != d->H || d->w != d->W) return false;
  const float* lse = (d->flags & B200SEG_WANT_LSE) ? d->lse : nullptr;
  const float* lpx = (d->flags & B200SEG_WANT_LOSS_PX) ? d->loss_px : nullptr;
  if ((lse && !aligned16(lse)) || (lpx && !aligned16(lpx))) return false;
  return bulk_supported(d->logits, d->labels, d->logits, d->logit_dtype, d->label_dtype, d->C, (long long)d->H * d->W,
                        d->pixel_weight != nullptr);
}

int bulk_fwd_dispatch(const b200seg_loss_desc* d, cudaStream_t st) {
  BulkParams p = {};
  p.logits = d->logits; p.labels = d->labels; p.cw = d->ce_class_weight;
  p.stats = reinterpret_cast<unsigned long long*>(d->stats);
  p.ce_scale_host = 0.f;
  p.lse = (d->flags & B200SEG_WANT_LSE) ? d->lse : nullptr;
  p.loss_px = (d->flags & B200SEG_WANT_LOSS_PX) ? d->loss_px : nullptr;
  p.lw = d->ce_loss_weight;
  p.label_dtype = d->label_dtype; p.label_bytes = label_bytes(d->label_dtype);
  p.N = d->N; p.C = d->C; p.HW = (long long)d->H * d->W;
  p.tiles_per_image = (int)((p.HW + kBulkPx - 1) / kBulkPx);
  p.total_tiles = (long long)p.tiles_per_image * p.N;
  p.ignore_index = d->ignore_index; p.acc_has_ignore = d->acc_has_ignore; p.acc_ignore = d->acc_ignore_index;
  switch (d->logit_dtype) {
    case B200SEG_F32: return bulk_launch<float, false>(p, st);
    case B200SEG_BF16: return bulk_launch<__nv_bfloat16, false>(p, st);
    case B200SEG_F16: return bulk_launch<__half, false>(p, st);
  }
  set_error("loss_fwd: unsupported logit dtype %d", d->logit_dtype);
  return 1;
}

}  // namespace b200seg
