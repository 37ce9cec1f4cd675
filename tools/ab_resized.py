#!/usr/bin/env python
"""Times the fused validation rescale + arg-max + areas kernel (row f2): 1/8-resolution logits against 1024x2048 ground truth."""
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
GH, GW = (int(v) for v in os.environ.get('AB_GT', '1024x2048').split('x'))
gt_base = [bench.make_labels((1, GH, GW), Cn, 800 + i, 255, dtype=torch.float32, device=dev)[0] for i in range(8)]
gts = [gt_base[i % 8].clone() for i in range(n)]
only = os.environ.get('AB_ONLY')
for name, shp, dt in (('1/8 fp32', (1, Cn, 128, 256), torch.float32), ('1/8 bf16', (1, Cn, 128, 256), torch.bfloat16),
                      ('1/4 fp32', (1, Cn, 256, 512), torch.float32), ('x3.41 fp32', (1, Cn, 300, 600), torch.float32),
                      ('x1.7 fp32', (1, Cn, 600, 1200), torch.float32), ('x2 fp32', (1, Cn, 512, 1024), torch.float32),
                      ('x0.8 fp32', (1, Cn, 1280, 2560), torch.float32)):
    if only and only != name:
        continue
    lo = [bench.make_logits(shp, 950 + i, dtype=dt, device=dev) for i in range(8)]
    lows = [lo[i % 8].clone() for i in range(n)]
    for ac in (False, True):
        tab = B.prepare_images(lows, gts, Cn, from_logits=True)

        def sweep(i):
            B.area_totals_device(tab, None, Cn, 255, align_corners=ac)

        sweep(0)
        ms = bench.timed_events(sweep, 3)
        print('%-11s ac=%d  %d images: %.3f ms  %.1f Gpix/s' % (name, ac, n, ms, n * GH * GW / ms / 1e6))
