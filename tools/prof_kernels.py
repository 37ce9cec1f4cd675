#!/usr/bin/env python
"""Small, fixed invocation of every hot kernel — the command line profiled with ncu (profiles/README.md).

    python tools/prof_kernels.py [c2] [c3] [c4] [c5] [lovasz]      (default: all)

Sizes are the BASELINE configs with reduced batch so that ncu's ~40 replays per kernel stay short; the
per-launch geometry (tile shapes, classes, dtypes) is that of the full configs.
"""
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import image_segmentation_lab_b200 as B  # noqa: E402


def main():
    warnings.simplefilter('ignore')
    which = set(sys.argv[1:]) or {'c2', 'c3', 'c4', 'c5', 'lovasz'}
    dev = torch.device('cuda', 0)
    reps = int(os.environ.get('PROF_REPS', '2'))
    if 'c2' in which:
        x = bench.make_logits((8, 19, 64, 128), 1, device=dev).requires_grad_(True)
        y = bench.make_labels((8, 512, 1024), 19, 1, device=dev).unsqueeze(1)
        ce = B.CrossEntropyLoss()
        for _ in range(reps):
            x.grad = None
            B.fused_resize_losses(x, y, ce, ignore_index=255)['loss_ce'].backward()
    if 'c2' in which:   # round 2: the thread-per-cell kernels — align_corners=True, and 150 classes (class-tiled plan)
        ce = B.CrossEntropyLoss()
        for _ in range(reps):
            x.grad = None
            B.fused_resize_losses(x, y, ce, align_corners=True, ignore_index=255)['loss_ce'].backward()
        x150 = bench.make_logits((2, 150, 64, 64), 1, device=dev).requires_grad_(True)
        y150 = bench.make_labels((2, 512, 512), 150, 1, device=dev).unsqueeze(1)
        for _ in range(reps):
            x150.grad = None
            B.fused_resize_losses(x150, y150, ce, ignore_index=255)['loss_ce'].backward()
    if 'c3' in which:
        x = bench.make_logits((4, 150, 512, 512), 2, dtype=torch.bfloat16, device=dev).requires_grad_(True)
        y = bench.make_labels((4, 512, 512), 150, 2, device=dev).unsqueeze(1)
        losses = [B.CrossEntropyLoss(class_weight=torch.linspace(0.5, 1.5, 150).tolist()), B.DiceLoss(loss_weight=3.0)]
        for _ in range(reps):
            x.grad = None
            r = B.fused_resize_losses(x, y, losses, ignore_index=255)
            (r['loss_ce'] + r['loss_dice']).backward()
    if 'c4' in which:
        x = bench.make_logits((8, 21, 512, 512), 3, device=dev).requires_grad_(True)
        y = bench.make_labels((8, 512, 512), 21, 3, device=dev).unsqueeze(1)
        for single in (True, False):
            ce = B.CrossEntropyLoss()
            ce.single_pass = single
            for _ in range(reps):
                x.grad = None
                B.fused_resize_losses(x, y, ce, ignore_index=255)['loss_ce'].backward()
    if 'c5' in which:
        g = torch.Generator(device=dev).manual_seed(5)
        n = 24
        gts = [bench.make_labels((1, 1024, 2048), 19, 50 + i, device=dev)[0].float() for i in range(n)]
        preds = [torch.randint(0, 19, (1024, 2048), generator=g, device=dev) for _ in range(n)]
        for _ in range(reps):
            B.areas_device(preds, gts, 19, 255)
        logits = [bench.make_logits((1, 19, 1024, 2048), 60 + i, device=dev) for i in range(8)]
        for _ in range(reps):
            B.areas_device(logits, gts[:8], 19, 255, from_logits=True)
        lows = [bench.make_logits((1, 19, 128, 256), 70 + i, device=dev) for i in range(8)]   # f2: rescale fused into the arg-max
        for _ in range(reps):
            B.areas_device(lows, gts[:8], 19, 255, from_logits=True)
    if 'lovasz' in which:   # Lovasz-Softmax at the config-2 label resolution, 4 classes' worth of segments
        x = bench.make_logits((8, 4, 512, 1024), 4, device=dev).requires_grad_(True)
        y = bench.make_labels((8, 512, 1024), 4, 4, device=dev)
        lv = B.LovaszLoss(reduction='none')
        for _ in range(reps):
            x.grad = None
            lv(x, y, ignore_index=255).backward()
    torch.cuda.synchronize()
    print('prof_kernels ok, launches =', B.launch_count())


if __name__ == '__main__':
    main()
