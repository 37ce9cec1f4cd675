// Micro-benchmark: per-SM throughput of FFMA, FFMA2 (fma.rn.f32x2), MUFU.EX2 and mixes on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_rates issue_rates.cu && ./issue_rates
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
  float a[8]; unsigned long long p[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; p[i] = (unsigned long long)__float_as_uint(a[i]) * 0x100000001ull; }
  const float m = 0.999f; const unsigned long long m2 = (unsigned long long)__float_as_uint(m) * 0x100000001ull;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) a[u] = fma1(a[u], m, m);                                    // 8 FFMA
      if (MODE == 1) p[u] = fma2(p[u], m2, m2);                                  // 8 FFMA2
      if (MODE == 2) a[u] = ex2(a[u]);                                           // 8 MUFU
      if (MODE == 3) { a[u] = fma1(a[u], m, m); if ((u & 3) == 0) a[u] = ex2(a[u]); }   // 8 FFMA + 2 MUFU
      if (MODE == 4) { p[u] = fma2(p[u], m2, m2); a[u] = fma1(a[u], m, m); }     // 8 FFMA2 + 8 FFMA
      if (MODE == 5) { p[u] = fma2(p[u], m2, m2); if ((u & 1) == 0) a[u] = ex2(a[u]); } // 8 FFMA2 + 4 MUFU
      if (MODE == 6) { p[u] = fma2(p[u], m2, m2); a[u] = __int_as_float(__float_as_int(a[u]) + u); } // 8 FFMA2 + 8 IADD
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, double inst_per_iter) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000; cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  k<MODE><<<148 * 8, 256>>>(out, 100, 1.f);
  cudaEventRecord(s); k<MODE><<<148 * 8, 256>>>(out, iters, 1.f); cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  const double warps = 148.0 * 8 * 8, winst = warps * iters * inst_per_iter;
  printf("%-28s %.3f ms  %.2f warp-inst/clk/SM (at 1.965 GHz)\n", name, ms, winst / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<0>("FFMA", 8); run<1>("FFMA2", 8); run<2>("MUFU.EX2", 8); run<3>("8 FFMA + 2 MUFU", 10);
  run<4>("8 FFMA2 + 8 FFMA", 16); run<5>("8 FFMA2 + 4 MUFU", 12); run<6>("8 FFMA2 + 8 IADD", 16);
  return 0;
}
