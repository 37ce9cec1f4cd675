#!/usr/bin/env python
"""Times the streaming kernels that request later lines into L2: ce_fwd (C3 shape, CE only, forward), bce (F1 shape), C5(i)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import image_segmentation_lab_b200 as B  # noqa: E402

dev = torch.device('cuda', 0)
peak, _ = bench.hbm_peak()


def run(name, shape, dtype, loss, algo_f, algo_fb, lab_dtype=torch.int64):
    xs = [bench.make_logits(shape, 300 + i, dtype=dtype, device=dev).requires_grad_(True) for i in range(2)]
    ys = [bench.make_labels((shape[0],) + shape[2:], shape[1], 300 + i, 255, device=dev).to(lab_dtype).unsqueeze(1) for i in range(2)]

    def fwd(i):
        with torch.no_grad():
            B.fused_resize_losses(xs[i & 1], ys[i & 1], loss, ignore_index=255)

    def fb(i):
        x = xs[i & 1]
        x.grad = None
        B.fused_resize_losses(x, ys[i & 1], loss, ignore_index=255)['loss_ce'].backward()

    for f, algo, tag in ((fwd, algo_f, 'fwd'), (fb, algo_fb, 'fwd+bwd')):
        for i in range(3):
            f(i)
        ms = min(bench.timed_events(f, 10) for _ in range(3))
        print('%-28s %-8s %.3f ms  frac %.3f' % (name, tag, ms, algo / (ms * 1e-3) / 1e9 / peak))


el3 = 16 * 150 * 512 * 512 * 2
px3 = 16 * 512 * 512
run('C3 bf16 CE only (two-pass)', (16, 150, 512, 512), torch.bfloat16, B.CrossEntropyLoss(), el3 + px3 * 8, 3 * el3 + 2 * px3 * 8)
el1 = 32 * 2 * 512 * 512 * 4
px1 = 32 * 512 * 512
run('F1 sigmoid CE', (32, 2, 512, 512), torch.float32, B.CrossEntropyLoss(use_sigmoid=True), el1 + px1 * 8, 2 * el1 + px1 * 8)
ce2 = B.CrossEntropyLoss()
ce2.single_pass = False
el4 = 32 * 21 * 512 * 512 * 4
run('C4 fp32 CE two-pass', (32, 21, 512, 512), torch.float32, ce2, el4 + px1 * 8, 3 * el4 + 2 * px1 * 8)
