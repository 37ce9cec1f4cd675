#!/usr/bin/env python
"""Times the label-map confusion sweep (BASELINE config 5 (i)): int64 predictions + float32 ground truth, 1024x2048."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import image_segmentation_lab_b200 as B  # noqa: E402

dev = torch.device('cuda', 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
Cn = 19
peak, _ = bench.hbm_peak()
gt_base = [bench.make_labels((1, 1024, 2048), Cn, 800 + i, 255, dtype=torch.float32, device=dev)[0] for i in range(8)]
g = torch.Generator(device=dev).manual_seed(5)
pred_base = [torch.randint(0, Cn, (1024, 2048), generator=g, device=dev) for _ in range(8)]
gts = [gt_base[i % 8].clone() for i in range(n)]
preds = [pred_base[i % 8].clone() for i in range(n)]
tab = B.prepare_images(preds, gts, Cn)


def sweep(i):
    B.area_totals_device(tab, None, Cn, 255)


sweep(0)
for rep in range(3):
    ms = bench.timed_events(sweep, 5)
    px = n * 1024 * 2048
    print('C5i %d images: %.3f ms  %.1f Gpix/s  frac %.3f' % (n, ms, px / ms / 1e6, px * 12 / (ms * 1e-3) / 1e9 / peak))
