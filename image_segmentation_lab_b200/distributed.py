"""Multi-GPU plumbing for the hot path: shard by image, ONE packed all-reduce.

Images are independent units: cross-entropy numerators, valid / correct pixel counts and the area
histograms are plain sums over pixels, so each rank (one process per GPU, torch.distributed over NCCL /
NVLink) works on its own contiguous image range with no data-path collective. The only exchange is one
all_reduce(SUM) of a small packed buffer — per step for the logged loss scalars (off the critical path: in
the default 'mean' mode the gradient's denominator N_global*H*W is known a priori), and once at the end of
an evaluation sweep for the area totals. The reference's counterpart is parse_losses'
per-variable all_reduce + .item() (utils/train_utils.py:56-72), which is never reached because train.py
does not initialise a process group.

Counts are packed as float64: every integer below 2**53 is exact, so summed areas stay bit-exact.
"""
import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib


def init_from_env(backend: Optional[str] = None, device: Optional[torch.device] = None):
    """One process per GPU as launched by torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kwargs = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kwargs['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local, world


def bind_to_gpu_numa_node(device_index: int) -> bool:
    """Pin this process to the CPUs NVML reports as local to the GPU, so that pinned host buffers allocated afterwards
    (first touch) and the copy threads sit on the GPU's NUMA node: with one process per GPU all ranks otherwise share one
    socket's memory path for their host-to-device traffic. Returns False (and changes nothing) when NVML or
    sched_setaffinity is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        finally:
            pynvml.nvmlShutdown()
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous ceil split: rank r owns [r*ceil(n/w), min(n, (r+1)*ceil(n/w)))."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


class PackedAllReduce:
    """Packs several small tensors into one float64 buffer and issues a single all_reduce(SUM).

    ``start`` enqueues the collective on a side stream ordered after the producer stream (so the compute
    stream never waits for NCCL); ``finish`` makes the current stream wait and returns the reduced
    tensors in their original dtypes and shapes. Without a process group both are identity operations.
    """

    def __init__(self, group=None, side_stream: bool = True):
        self.group = group
        self._side = None
        self._want_side = side_stream
        self._inflight = None

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def start(self, tensors: Sequence[torch.Tensor]):
        assert self._inflight is None, 'previous all-reduce not finished'
        metas = [(t.dtype, tuple(t.shape), t.numel()) for t in tensors]
        if self._world() == 1:
            self._inflight = (None, None, metas, list(tensors))
            return
        flat = torch.cat([t.reshape(-1).to(torch.float64) for t in tensors])
        work = None
        if flat.is_cuda and self._want_side:
            if self._side is None:
                self._side = torch.cuda.Stream(device=flat.device)
            self._side.wait_stream(torch.cuda.current_stream(flat.device))
            with torch.cuda.stream(self._side):
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.record_stream(self._side)
        else:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight = (flat, work, metas, None)

    def finish(self) -> List[torch.Tensor]:
        flat, work, metas, passthrough = self._inflight
        self._inflight = None
        if passthrough is not None:
            return passthrough
        if work is not None:
            work.wait()
        elif flat.is_cuda and self._side is not None:
            torch.cuda.current_stream(flat.device).wait_stream(self._side)
        out, off = [], 0
        for dtype, shape, numel in metas:
            piece = flat[off:off + numel]
            off += numel
            if not dtype.is_floating_point:
                piece = piece.round()
            out.append(piece.to(dtype).reshape(shape))
        return out

    def __call__(self, tensors):
        self.start(tensors)
        return self.finish()


def all_reduce_areas(areas: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """Sum int64 area tensors over ranks with ONE collective for all keys (sorted, so ranks agree). NCCL and gloo sum
    int64 natively, so the counts travel as they are: no float64 packing, no rounding pass, and for a single key not
    even a concatenation — the reduction happens in place on the stream the areas were produced on."""
    keys = sorted(areas.keys())
    if not keys:
        return {}
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return {k: areas[k] for k in keys}
    for k in keys:
        assert areas[k].dtype == torch.int64, 'area tensors are int64 (got %s for %r)' % (areas[k].dtype, k)
    if len(keys) == 1:
        t = areas[keys[0]].contiguous()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return {keys[0]: t}
    flat = torch.cat([areas[k].reshape(-1) for k in keys])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, off = {}, 0
    for k in keys:
        n = areas[k].numel()
        out[k] = flat[off:off + n].reshape(areas[k].shape)
        off += n
    return out


def global_loss_scalars(vec: torch.Tensor, loss_weight: float = 1.0, avg_non_ignore: bool = False):
    """Global-batch CE loss / accuracy from the all-reduced statistics vector of the fused loss
    (``fused_resize_losses(..., return_stats=True)['_stats']``, layout B200SEG_LOG_*): identical to running the
    reference once on the concatenated batch — loss = lw * sum / (N_global*H*W), acc = 100 * correct / n_acc."""
    eps = float(torch.finfo(torch.float32).eps)
    denom = (vec[_lib.LOG_N_VALID] + eps) if avg_non_ignore else vec[_lib.LOG_N_PIXELS]
    loss = (loss_weight * vec[_lib.LOG_CE_SUM] / denom).to(torch.float32)
    acc = (100.0 * (vec[_lib.LOG_N_CORRECT] + eps) / (vec[_lib.LOG_N_ACC] + eps)).to(torch.float32)
    return loss, acc


def global_dice_loss(vec: torch.Tensor, num_classes: int, loss_weight: float = 1.0):
    """Global-batch Dice / Tversky loss from the all-reduced statistics vector: the per-(sample, class) terms are additive
    over images (dice_loss.py:31-58: class mean of the batch means), so loss = lw * sum / (C * N_global)."""
    return (loss_weight * vec[_lib.LOG_DICE_SUM] / (float(num_classes) * vec[_lib.LOG_N_IMAGES])).to(torch.float32)
