// Shared device helpers for libb200seg (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200seg.h"

namespace b200seg {

constexpr float kLog2e = 1.4426950408889634f;
constexpr int kSMs = 148;                            // B200: 2 dies x 74 SMs

// ---------------------------------------------------------------- host-side plumbing
void set_error(const char* fmt, ...);
int check_launch(const char* what);
void count_launch(int n = 1);
int ensure_dyn_smem(const void* func, int bytes);   // per (kernel, device) opt-in to > 48 KB of dynamic shared memory

#define B200SEG_REQUIRE(cond, ...)       \
  do {                                   \
    if (!(cond)) {                       \
      ::b200seg::set_error(__VA_ARGS__); \
      return 1;                          \
    }                                    \
  } while (0)

#define B200SEG_CUDA(expr)                                                         \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      ::b200seg::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));        \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
__host__ __device__ inline int label_bytes(int dt) {
  switch (dt) {
    case B200SEG_L_U8: return 1;
    case B200SEG_L_I16: return 2;
    case B200SEG_L_I32: return 4;
    case B200SEG_L_I64: return 8;
    case B200SEG_L_F32: return 4;
    case B200SEG_L_F64: return 8;
  }
  return 0;
}
inline int logit_bytes(int dt) { return dt == B200SEG_F32 ? 4 : 2; }

// ---------------------------------------------------------------- math
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

// ---------------------------------------------------------------- streaming loads / stores
// Logits are read exactly once: bypass L1 allocation so the L1 keeps the small tables.
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_stream8(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream4(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream2(const void* p) {
  uint16_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream1(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream8(void* p, uint2 v) {
  asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream4(void* p, uint32_t v) {
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream2(void* p, uint32_t v) {
  asm volatile("st.global.cs.u16 [%0], %1;" ::"l"(p), "h"((uint16_t)v) : "memory");
}

// ---------------------------------------------------------------- logit element conversion
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int kBytes = 4;
  static constexpr int kDtype = B200SEG_F32;
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int kBytes = 2;
  static constexpr int kDtype = B200SEG_BF16;
};
template <> struct Elem<__half> {
  static constexpr int kBytes = 2;
  static constexpr int kDtype = B200SEG_F16;
};

template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }

// two packed 16-bit elements <-> two floats
template <typename T> __device__ __forceinline__ void unpack2(uint32_t r, float& lo, float& hi);
template <> __device__ __forceinline__ void unpack2<__nv_bfloat16>(uint32_t r, float& lo, float& hi) {
  lo = __uint_as_float(r << 16);
  hi = __uint_as_float(r & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack2<__half>(uint32_t r, float& lo, float& hi) {
  __half2 h = *reinterpret_cast<__half2*>(&r);
  float2 f = __half22float2(h);
  lo = f.x;
  hi = f.y;
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Load V consecutive elements of T (V*sizeof(T) in {2,4,8,16} bytes, pointer aligned to that) as floats.
template <typename T, int V> __device__ __forceinline__ void load_vec(const T* p, float (&o)[V]) {
  if constexpr (sizeof(T) == 4) {
    if constexpr (V == 4) {
      uint4 r = ld_stream16(p);
      o[0] = __uint_as_float(r.x); o[1] = __uint_as_float(r.y);
      o[2] = __uint_as_float(r.z); o[3] = __uint_as_float(r.w);
    } else if constexpr (V == 2) {
      uint2 r = ld_stream8(p);
      o[0] = __uint_as_float(r.x); o[1] = __uint_as_float(r.y);
    } else {
      static_assert(V == 1, "fp32: V in {1,2,4}");
      o[0] = __uint_as_float(ld_stream4(p));
    }
  } else {
    if constexpr (V == 8) {
      uint4 r = ld_stream16(p);
      unpack2<T>(r.x, o[0], o[1]); unpack2<T>(r.y, o[2], o[3]);
      unpack2<T>(r.z, o[4], o[5]); unpack2<T>(r.w, o[6], o[7]);
    } else if constexpr (V == 4) {
      uint2 r = ld_stream8(p);
      unpack2<T>(r.x, o[0], o[1]); unpack2<T>(r.y, o[2], o[3]);
    } else if constexpr (V == 2) {
      unpack2<T>(ld_stream4(p), o[0], o[1]);
    } else {
      static_assert(V == 1, "16-bit: V in {1,2,4,8}");
      float hi;
      unpack2<T>(ld_stream2(p), o[0], hi);
    }
  }
}

// V consecutive fp32 values of a per-pixel map (V up to 8: two 128-bit accesses)
template <int V> __device__ __forceinline__ void load_px_f32(const float* p, float (&o)[V]) {
  if constexpr (V == 8) {
    float a[4], b[4];
    load_vec<float, 4>(p, a);
    load_vec<float, 4>(p + 4, b);
#pragma unroll
    for (int k = 0; k < 4; ++k) { o[k] = a[k]; o[4 + k] = b[k]; }
  } else {
    load_vec<float, V>(p, o);
  }
}

// Raw (still packed) form of V consecutive elements: loads stay in as few registers as the bytes they carry, so a
// thread can keep more bytes in flight; convert with unpack_raw() only where the values are consumed.
template <typename T, int V> struct RawVec {
  static constexpr int W = (V * (int)sizeof(T) + 3) / 4;
  uint32_t w[W];
};
template <typename T, int V> __device__ __forceinline__ RawVec<T, V> load_raw(const T* p) {
  RawVec<T, V> r;
  constexpr int B = V * (int)sizeof(T);
  if constexpr (B == 16) { const uint4 t = ld_stream16(p); r.w[0] = t.x; r.w[1] = t.y; r.w[2] = t.z; r.w[3] = t.w; }
  else if constexpr (B == 8) { const uint2 t = ld_stream8(p); r.w[0] = t.x; r.w[1] = t.y; }
  else if constexpr (B == 4) { r.w[0] = ld_stream4(p); }
  else { static_assert(B == 2, "raw vector of 2, 4, 8 or 16 bytes"); r.w[0] = ld_stream2(p); }
  return r;
}
template <typename T, int V> __device__ __forceinline__ void unpack_raw(const RawVec<T, V>& r, float (&o)[V]) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int v = 0; v < V; ++v) o[v] = __uint_as_float(r.w[v]);
  } else if constexpr (V == 1) {
    float hi;
    unpack2<T>(r.w[0], o[0], hi);
  } else {
#pragma unroll
    for (int v = 0; v < V; v += 2) unpack2<T>(r.w[v / 2], o[v], o[v + 1]);
  }
}

template <typename T, int V> __device__ __forceinline__ void store_vec(T* p, const float (&v)[V]) {
  if constexpr (sizeof(T) == 4) {
    if constexpr (V == 4) {
      st_stream16(p, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
    } else if constexpr (V == 2) {
      st_stream8(p, make_uint2(__float_as_uint(v[0]), __float_as_uint(v[1])));
    } else {
      st_stream4(p, __float_as_uint(v[0]));
    }
  } else {
    if constexpr (V == 8) {
      st_stream16(p, make_uint4(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]), pack2<T>(v[4], v[5]), pack2<T>(v[6], v[7])));
    } else if constexpr (V == 4) {
      st_stream8(p, make_uint2(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3])));
    } else if constexpr (V == 2) {
      st_stream4(p, pack2<T>(v[0], v[1]));
    } else {
      st_stream2(p, pack2<T>(v[0], 0.f) & 0xffffu);
    }
  }
}

// ---------------------------------------------------------------- labels
// Labels are consumed in the dtype the data pipeline delivers (reference: label.long(),
// cross_entropy_loss.py:283; evaluator ground truth is float32, core/dataset/kvasir_seg.py:37).
__device__ __forceinline__ long long load_label(const void* p, int dt, size_t i) {
  switch (dt) {
    case B200SEG_L_U8: return (long long)((const uint8_t*)p)[i];
    case B200SEG_L_I16: return (long long)((const int16_t*)p)[i];
    case B200SEG_L_I32: return (long long)((const int32_t*)p)[i];
    case B200SEG_L_I64: return ((const long long*)p)[i];
    case B200SEG_L_F32: return (long long)((const float*)p)[i];
    default: return (long long)((const double*)p)[i];
  }
}

// V consecutive labels starting at element i (i % V == 0 and base 16-byte aligned when V > 1).
template <int V> __device__ __forceinline__ void load_labels(const void* p, int dt, size_t i, long long (&o)[V]) {
  if constexpr (V == 1) {
    o[0] = load_label(p, dt, i);
  } else {
    static_assert(V == 2 || V == 4 || V == 8, "V");
    switch (dt) {
      case B200SEG_L_I64:
      case B200SEG_L_F64: {
        const uint4* q = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(p) + i * 8);
#pragma unroll
        for (int k = 0; k < V / 2; ++k) {
          uint4 r = ld_stream16(q + k);
          long long a = (long long)(((unsigned long long)r.y << 32) | r.x);
          long long b = (long long)(((unsigned long long)r.w << 32) | r.z);
          if (dt == B200SEG_L_F64) {
            a = (long long)__longlong_as_double(a);
            b = (long long)__longlong_as_double(b);
          }
          o[2 * k] = a;
          o[2 * k + 1] = b;
        }
      } break;
      case B200SEG_L_I32:
      case B200SEG_L_F32: {
        uint32_t r[V];
        const char* q = reinterpret_cast<const char*>(p) + i * 4;
        if constexpr (V == 8) {
          uint4 a = ld_stream16(q), b = ld_stream16(q + 16);
          r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
        } else if constexpr (V == 4) {
          uint4 a = ld_stream16(q);
          r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
        } else {
          uint2 a = ld_stream8(q);
          r[0] = a.x; r[1] = a.y;
        }
#pragma unroll
        for (int k = 0; k < V; ++k)
          o[k] = (dt == B200SEG_L_F32) ? (long long)__uint_as_float(r[k]) : (long long)(int32_t)r[k];
      } break;
      case B200SEG_L_I16: {
        const int16_t* q = reinterpret_cast<const int16_t*>(p) + i;
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] = (long long)q[k];
      } break;
      default: {  // U8
        const char* q = reinterpret_cast<const char*>(p) + i;
        uint32_t r0, r1 = 0;
        if constexpr (V == 8) {
          uint2 a = ld_stream8(q);
          r0 = a.x; r1 = a.y;
        } else if constexpr (V == 4) {
          r0 = ld_stream4(q);
        } else {
          r0 = ld_stream2(q);
        }
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] = (long long)(((k < 4 ? r0 : r1) >> (8 * (k & 3))) & 0xffu);
      } break;
    }
  }
}

// ---------------------------------------------------------------- programmatic dependent launch (sm_90+)
// A kernel launched with launch_pdl() may be scheduled while its predecessor in the stream is still draining; it must
// call pdl_wait() before touching anything the predecessor wrote (no-op for a plain launch). A predecessor that calls
// pdl_launch_dependents() lets that scheduling begin early; without it the dependents start when it exits.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------- reductions
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of up to NV values per thread; result valid in thread 0. `scratch` holds
// NV * 32 elements of T. Requires blockDim.x to be a multiple of 32.
template <typename T, int NV> __device__ __forceinline__ void block_sum(T (&v)[NV], T* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) scratch[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      T x = lane < nwarp ? scratch[k * 32 + lane] : T(0);
      v[k] = warp_sum(x);
    }
  }
}

// CTA-wide flush of the per-thread loss / counter partials into the 64-bit statistics block: one REDUX per
// counter and one shuffle tree for the loss per warp, one shared-memory hop, then <= 5 global atomics per CTA.
// (The generic block_sum<double,5> costs ~150 instructions per thread; this one ~25.)
__device__ __forceinline__ void cta_flush_stats(float loss, int n_valid, int n_correct, int n_bad, int n_acc,
                                                unsigned long long* stats, bool want_ce = true) {
  __shared__ float s_loss[32];
  __shared__ int s_cnt[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  loss = warp_sum(loss);
  n_valid = __reduce_add_sync(0xffffffffu, n_valid);
  n_correct = __reduce_add_sync(0xffffffffu, n_correct);
  n_bad = __reduce_add_sync(0xffffffffu, n_bad);
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  if (lane == 0) {
    s_loss[warp] = loss;
    s_cnt[0][warp] = n_valid; s_cnt[1][warp] = n_correct; s_cnt[2][warp] = n_bad; s_cnt[3][warp] = n_acc;
  }
  __syncthreads();
  if (warp == 0) {
    double l = lane < nwarp ? (double)s_loss[lane] : 0.0;
    l = warp_sum(l);
    const int a = __reduce_add_sync(0xffffffffu, lane < nwarp ? s_cnt[0][lane] : 0);
    const int b = __reduce_add_sync(0xffffffffu, lane < nwarp ? s_cnt[1][lane] : 0);
    const int c = __reduce_add_sync(0xffffffffu, lane < nwarp ? s_cnt[2][lane] : 0);
    const int d = __reduce_add_sync(0xffffffffu, lane < nwarp ? s_cnt[3][lane] : 0);
    if (lane == 0) {
      if (want_ce) {
        atomicAdd(reinterpret_cast<double*>(stats + B200SEG_ST_CE_SUM), l);
        atomicAdd(stats + B200SEG_ST_N_VALID, (unsigned long long)a);
        if (c) atomicAdd(stats + B200SEG_ST_N_BAD, (unsigned long long)c);
      }
      atomicAdd(stats + B200SEG_ST_N_CORRECT, (unsigned long long)b);
      atomicAdd(stats + B200SEG_ST_N_ACC, (unsigned long long)d);
    }
  }
}

// The same for a CTA of ONE warp: no shared-memory hop, no barrier.
__device__ __forceinline__ void warp_flush_stats(float loss, int n_valid, int n_correct, int n_bad, int n_acc,
                                                 unsigned long long* stats) {
  loss = warp_sum(loss);
  n_valid = __reduce_add_sync(0xffffffffu, n_valid);
  n_correct = __reduce_add_sync(0xffffffffu, n_correct);
  n_bad = __reduce_add_sync(0xffffffffu, n_bad);
  n_acc = __reduce_add_sync(0xffffffffu, n_acc);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(reinterpret_cast<double*>(stats + B200SEG_ST_CE_SUM), (double)loss);
    atomicAdd(stats + B200SEG_ST_N_VALID, (unsigned long long)n_valid);
    if (n_bad) atomicAdd(stats + B200SEG_ST_N_BAD, (unsigned long long)n_bad);
    atomicAdd(stats + B200SEG_ST_N_CORRECT, (unsigned long long)n_correct);
    atomicAdd(stats + B200SEG_ST_N_ACC, (unsigned long long)n_acc);
  }
}

// One-hot Dice sums of one pixel per lane into this warp's class bins: lanes holding the same class are found with one
// MATCH.ANY, their p_y are added with one integer REDUX in Q23 fixed point (p_y <= 1, 32 lanes: no overflow; rounding
// 6e-8 per term, unbiased), and the group's lowest lane does the two shared-memory updates. cls < 0 = nothing to add.
__device__ __forceinline__ void onehot_bins_add(float* A_w, float* T_w, int cls, float pyv, int lane) {
  const unsigned peers = __match_any_sync(0xffffffffu, cls);
  const unsigned tot = __reduce_add_sync(peers, (unsigned)__float2int_rn(pyv * 8388608.f));
  if (cls >= 0 && lane == __ffs(peers) - 1) {
    A_w[cls] += (float)tot * (1.f / 8388608.f);
    T_w[cls] += (float)__popc(peers);
  }
  __syncwarp();
}

// lg2.approx / rcp.approx: 1 MUFU each (max rel. error 2^-22 / 1 ulp) — used once per pixel
__device__ __forceinline__ float fast_log(float x) { return __log2f(x) * 0.6931471805599453f; }
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------- bilinear resize, ATen's arithmetic operation for operation
// torch/include/ATen/native/UpSample.h:271-312 (area_pixel_compute_scale / _source_index) and the expression of
// upsample_bilinear2d_out_frame, with the FMA contraction nvcc gives ATen's build PINNED by intrinsics: measured with
// tools/probe/run_fma_probe.py on B200 / torch 2.11.0+cu128 — of 29 candidate evaluations exactly one reproduces
// F.interpolate bit for bit (0 of 9.6 M elements differ, both align_corners settings, non-power-of-two shapes):
//     src = fma(scale, dst + 0.5, -0.5)                   (align_corners=False; scale * dst otherwise)
//     X   = fma(w0, v00, w1 * v01),  Y = fma(w0, v10, w1 * v11),  out = fma(h0, X, h1 * Y)
__host__ __device__ __forceinline__ float resize_scale(int in, int out, bool align_corners) {
  if (align_corners) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  return (float)in / (float)out;
}
__device__ __forceinline__ float aten_src_index(float scale, int dst, bool align_corners) {
  if (align_corners) return __fmul_rn(scale, (float)dst);
  const float s = __fmaf_rn(scale, __fadd_rn((float)dst, 0.5f), -0.5f);
  return s < 0.f ? 0.f : s;
}
__device__ __forceinline__ float aten_bilerp(float h0, float h1, float w0, float w1, float v00, float v01, float v10, float v11) {
  const float X = __fmaf_rn(w0, v00, __fmul_rn(w1, v01));
  const float Y = __fmaf_rn(w0, v10, __fmul_rn(w1, v11));
  return __fmaf_rn(h0, X, __fmul_rn(h1, Y));
}
// taps (i0, i1) and the weight l1 of tap i1 (the weight of i0 is 1 - l1, formed by the caller as ATen does: 1.f - l1)
__device__ __forceinline__ void resize_src(float scale, int dst, int in, bool align_corners, int& i0, int& i1,
                                           float& l1) {
  const float src = aten_src_index(scale, dst, align_corners);
  int i = (int)src;
  i = i < in - 1 ? i : in - 1;      // never active for a valid source index (< in); keeps the taps in range regardless
  i0 = i;
  i1 = i + (i < in - 1 ? 1 : 0);
  l1 = __fsub_rn(src, (float)i);
}

}  // namespace b200seg
