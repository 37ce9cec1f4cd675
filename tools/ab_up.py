#!/usr/bin/env python
"""Times the resize-fused CE forward+backward kernel alone through the C ABI (CUDA events), for A/B runs:

    B200SEG_UP_IMPL=band python tools/ab_up.py ; python tools/ab_up.py ; B200SEG_UPCELL_RG=2 python tools/ab_up.py
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import image_segmentation_lab_b200 as B  # noqa: E402
from image_segmentation_lab_b200 import _lib  # noqa: E402


def run(N, Cc, h, w, S, dtype=torch.float32, ldt=torch.int64, iters=100):
    dev = torch.device('cuda', 0)
    lib = B.load_library()
    H, W = h * S, w * S
    R = 6
    xs = [bench.make_logits((N, Cc, h, w), 10 + i, dtype=dtype, device=dev) for i in range(R)]
    ys = [bench.make_labels((N, H, W), Cc, 10 + i, 255, device=dev, dtype=ldt) for i in range(R)]
    nbytes = lib.b200seg_loss_fused_workspace_bytes(N, Cc, h, w, H, W, 0)
    pb = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    descs = []
    for i in range(R):
        fu = _lib.LossFusedDesc()
        fd = fu.fwd
        fd.logits = xs[i].data_ptr(); fd.labels = ys[i].data_ptr()
        fd.logit_dtype = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}[dtype]
        fd.label_dtype = {torch.int64: _lib.L_I64, torch.uint8: _lib.L_U8}[ldt]
        fd.N, fd.C, fd.h, fd.w, fd.H, fd.W = N, Cc, h, w, H, W
        fd.flags = _lib.WANT_CE | _lib.WANT_ACC
        fd.ignore_index = 255; fd.acc_has_ignore = 1; fd.acc_ignore_index = 255
        fd.dice_exponent = 2.0; fd.ce_loss_weight = 1.0
        fd.stats = stats.data_ptr()
        fu.grad_scale_host = 1.0
        fu.workspace = pb.data_ptr()
        fu.defer_combine = 1
        descs.append(fu)
    stream = _lib.stream_ptr(dev)

    def call(i):
        rc = lib.b200seg_loss_fused_fwdbwd(C.byref(descs[i % R]), stream)
        if rc:
            raise RuntimeError(_lib.last_error())

    for i in range(10):
        call(i)
    ms = bench.timed_events(call, iters)
    px = N * H * W
    print('impl=%s var=%s rg=%s  N%d C%d %dx%d S%d %s/%s: %.1f us  %.1f Gpix/s' % (
        os.environ.get('B200SEG_UP_IMPL', 'cell'), os.environ.get('B200SEG_UPCELL_VAR', '-'), os.environ.get('B200SEG_UPCELL_RG', 'auto'), N, Cc, h, w, S,
        str(dtype).replace('torch.', ''), str(ldt).replace('torch.', ''), ms * 1e3, px / ms / 1e6))


if __name__ == '__main__':
    run(8, 19, 64, 128, 8)
    if os.environ.get('AB_ONLY'):
        sys.exit(0)
    run(8, 19, 64, 128, 8, ldt=torch.uint8)
    run(8, 19, 64, 128, 8, dtype=torch.bfloat16)
    run(8, 19, 128, 256, 4)
    run(8, 19, 32, 64, 16)
    run(8, 8, 64, 128, 8)
    run(8, 32, 64, 128, 8)
    run(8, 19, 64, 128, 8, dtype=torch.float16)
