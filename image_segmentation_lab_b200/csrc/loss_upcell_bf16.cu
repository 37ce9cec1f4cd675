// Resize-fused cell-owner CE kernels, __nv_bfloat16 logits (see loss_upcell.cuh); one translation unit per dtype to compile in parallel.
#include "loss_upcell.cuh"

namespace b200seg {
template int upcell_run<__nv_bfloat16>(const b200seg_loss_desc*, float*, int, bool, cudaStream_t);
}
