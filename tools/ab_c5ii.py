#!/usr/bin/env python
"""Times the arg-max + areas sweep from full-resolution fp32 logits (BASELINE config 5 (ii))."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import image_segmentation_lab_b200 as B  # noqa: E402

dev = torch.device('cuda', 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
Cn = 19
peak, _ = bench.hbm_peak()
gt_base = [bench.make_labels((1, 1024, 2048), Cn, 800 + i, 255, dtype=torch.float32, device=dev)[0] for i in range(4)]
lbase = [bench.make_logits((1, Cn, 1024, 2048), 900 + i, device=dev) for i in range(4)]
gts = [gt_base[i % 4].clone() for i in range(n)]
logits = [lbase[i % 4].clone() for i in range(n)]
tab = B.prepare_images(logits, gts, Cn, from_logits=True)


def sweep(i):
    B.area_totals_device(tab, None, Cn, 255)


sweep(0)
for rep in range(3):
    ms = bench.timed_events(sweep, 3)
    px = n * 1024 * 2048
    print('C5ii %d images: %.3f ms  %.1f Gpix/s  frac %.3f' % (n, ms, px / ms / 1e6, px * (Cn * 4 + 4) / (ms * 1e-3) / 1e9 / peak))
