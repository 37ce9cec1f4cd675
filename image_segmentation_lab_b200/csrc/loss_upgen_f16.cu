// Resize-fused thread-per-cell CE kernels, __half logits (see loss_upgen.cuh); one translation unit per dtype to compile in parallel.
#include "loss_upgen.cuh"

namespace b200seg {
template int upgen_run<__half>(const b200seg_loss_desc*, float*, bool, cudaStream_t);
}
