// Register-tile kernels (loss_rt.cuh) instantiated for float logits — one translation unit per dtype so that the
// template instantiations build in parallel.
#include "loss_rt.cuh"

namespace b200seg {
int rt_run_f32(const RtParams& p, int kind, bool vec, cudaStream_t st) { return rt_run<float>(p, kind, vec, st); }
}  // namespace b200seg
