#!/usr/bin/env python
"""Times BASELINE config 4 (fp32, 21 classes, 512x512, batch 32) CE forward+backward, single-pass and two-pass plans."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import image_segmentation_lab_b200 as B  # noqa: E402

dev = torch.device('cuda', 0)
peak, _ = bench.hbm_peak()
shape = (32, 21, 512, 512)
xs = [bench.make_logits(shape, 300 + i, device=dev).requires_grad_(True) for i in range(2)]
ys = [bench.make_labels((32, 512, 512), 21, 300 + i, 255, device=dev).unsqueeze(1) for i in range(2)]
ce0 = B.CrossEntropyLoss()


def fwd_only(i):
    with torch.no_grad():
        B.fused_resize_losses(xs[i & 1], ys[i & 1], ce0, ignore_index=255)


for i in range(3):
    fwd_only(i)
ms = bench.timed_events(fwd_only, 20)
print('C4 fwd (no grad): %.3f ms  frac %.2f' % (ms, (32 * 21 * 512 * 512 * 4 + 32 * 512 * 512 * 8) / (ms * 1e-3) / 1e9 / peak))
for single in (True, False):
    ce = B.CrossEntropyLoss()
    ce.single_pass = single

    def fb(i):
        x = xs[i & 1]
        x.grad = None
        B.fused_resize_losses(x, ys[i & 1], ce, ignore_index=255)['loss_ce'].backward()

    for i in range(3):
        fb(i)
    ms = bench.timed_events(fb, 20)
    el = 32 * 21 * 512 * 512 * 4
    px = 32 * 512 * 512
    algo = (2 * el + px * 8) if single else (3 * el + 2 * px * 8)
    print('C4 %s MINB=%s: %.3f ms  frac %.2f' % ('single' if single else 'two-pass', os.environ.get('B200SEG_RT_MINB', '4'), ms,
                                                 algo / (ms * 1e-3) / 1e9 / peak))
