"""Worker of tests/test_gpu_round2.py::test_multi_gpu_nccl_sharded_totals_match_single_gpu — launched under torchrun with
one process per GPU (NCCL). Shards a batch and an image list by image (distributed.shard_range), runs the CUDA path on
the shard, all-reduces the additive statistics / int64 areas, and rank 0 saves what a single GPU must reproduce."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import image_segmentation_lab_b200 as B                      # noqa: E402
from image_segmentation_lab_b200 import distributed as D    # noqa: E402
from tests.helpers import synth_labels, synth_logits        # noqa: E402


def main():
    rank, local, world = D.init_from_env('nccl')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    N, C = 8, 21
    x_all = synth_logits((N, C, 128, 128), 77)
    y_all = synth_labels((N, 128, 128), C, 77, block=8)
    lo, hi = D.shard_range(N, rank, world)
    x = x_all[lo:hi].to(dev).requires_grad_(True)
    y = y_all[lo:hi].to(dev)
    r = B.fused_resize_losses(x, y.unsqueeze(1), [B.CrossEntropyLoss(), B.DiceLoss(loss_weight=3.0)], ignore_index=255,
                              return_stats=True)
    (r['loss_ce'] + r['loss_dice']).backward()
    vec = r['_stats'].clone()
    dist.all_reduce(vec, op=dist.ReduceOp.SUM)               # ONE all-reduce of the 8-double statistics vector
    loss_ce, acc = D.global_loss_scalars(vec, loss_weight=1.0)
    loss_dice = D.global_dice_loss(vec, C, loss_weight=3.0)
    # local 'mean' losses divide by the LOCAL pixel / image count: the global-batch gradient of an equal shard is 1/world of it
    g_local = x.grad * ((hi - lo) / float(N))
    g_all = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(g_all, g_local)

    n_img = 16
    lo, hi = D.shard_range(n_img, rank, world)
    preds = [torch.randint(0, 19, (96, 160), generator=torch.Generator().manual_seed(900 + i)).to(dev) for i in range(lo, hi)]
    gts = [synth_labels((1, 96, 160), 19, 900 + i, block=8)[0].float().to(dev) for i in range(lo, hi)]
    # parse_losses under NCCL: one all-reduce for every logged variable of the step, rank mean as the reference
    # (utils/train_utils.py:56-72)
    _, logged = B.parse_losses({'loss_ce': r['loss_ce'].detach(), 'loss_dice': r['loss_dice'].detach(), 'acc_seg': r['acc_seg']})
    tot = B.area_totals_device(preds, gts, 19, 255)
    tot = D.all_reduce_areas({'areas': tot})['areas']
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({'loss_ce': float(loss_ce), 'loss_dice': float(loss_dice), 'acc_seg': float(acc),
                    'grad': torch.cat(g_all, 0).cpu(), 'areas': tot.cpu(), 'logged': dict(logged),
                    'local_loss_ce': float(r['loss_ce'])}, os.environ['MGPU_OUT'])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
