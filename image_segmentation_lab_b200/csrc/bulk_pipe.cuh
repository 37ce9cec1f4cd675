// mbarrier / bulk-copy (TMA engine, non-tensor form) primitives shared by the bulk-copy pipelines, sm_100a.
// SASS: cp.async.bulk -> UBLKCP.S.G / UBLKCP.G.S, expect_tx -> SYNCS.ARRIVE.TRANS64.
#pragma once
#include "common.cuh"

namespace b200seg {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// suspend-time hint of try_wait: a waiting warp sleeps in hardware for up to this long instead of re-issuing the poll (the
// polls of waiting warps were 6 % of cs_fwd_kernel's issued instructions)
constexpr unsigned kMbarSuspendNs = 20000u;
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}

// Stage index / phase parity of a ring of NS stages walked in steps of `step` (no division per tile).
struct StageRing {
  int s, ph, ns, step;
  __device__ __forceinline__ void init(int first, int ns_, int step_) {
    ns = ns_; step = step_;
    s = first % ns_;
    ph = (first / ns_) & 1;
  }
  __device__ __forceinline__ void advance() {
    s += step;
    while (s >= ns) { s -= ns; ph ^= 1; }
  }
};

__device__ __forceinline__ long long smem_label(const unsigned char* row, int dt, int t) {
  if (dt == B200SEG_L_I64) return reinterpret_cast<const long long*>(row)[t];   // the two common cases first: one compare
  if (dt == B200SEG_L_U8) return (long long)row[t];
  switch (dt) {
    case B200SEG_L_U8: return (long long)row[t];
    case B200SEG_L_I16: return (long long)reinterpret_cast<const short*>(row)[t];
    case B200SEG_L_I32: return (long long)reinterpret_cast<const int*>(row)[t];
    case B200SEG_L_I64: return reinterpret_cast<const long long*>(row)[t];
    case B200SEG_L_F32: return (long long)reinterpret_cast<const float*>(row)[t];
    default: return (long long)reinterpret_cast<const double*>(row)[t];
  }
}

}  // namespace b200seg
