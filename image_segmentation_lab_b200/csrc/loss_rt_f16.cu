// Register-tile kernels (loss_rt.cuh) instantiated for __half logits — one translation unit per dtype so that the
// template instantiations build in parallel.
#include "loss_rt.cuh"

namespace b200seg {
int rt_run_f16(const RtParams& p, int kind, bool vec, cudaStream_t st) { return rt_run<__half>(p, kind, vec, st); }
}  // namespace b200seg
