#!/usr/bin/env python
"""Runs tools/probe/fma_probe.cu against F.interpolate(mode='bilinear') on this GPU: prints, per evaluation variant, how
many output elements differ bitwise from ATen (and the largest difference in ulps) over non-power-of-two shapes, both
align_corners settings."""
import ctypes
import os
import subprocess

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'libfma_probe.so')


def build():
    src = os.path.join(HERE, 'fma_probe.cu')
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-shared', '-Xcompiler', '-fPIC', '-cudart',
                        'static', src, '-o', SO], check=True)
    lib = ctypes.CDLL(SO)
    lib.probe_launch.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_float] * 2 + [ctypes.c_int] * 4 + [
        ctypes.c_void_p]
    lib.probe_launch.restype = ctypes.c_int
    return lib


def main():
    lib = build()
    if not torch.cuda.is_available():
        print('built', SO)
        return
    dev = torch.device('cuda', 0)
    shapes = [((3, 19, 47), (513, 1025)), ((2, 65, 129), (513, 1025)), ((5, 33, 31), (100, 77)), ((7, 60, 90), (61, 91)),
              ((3, 100, 100), (37, 53)), ((4, 64, 128), (512, 1024))]
    names = {0: 'fma(x,a,y*b)', 1: 'fma(y,b,x*a)', 2: 'no fma'}
    tot, worst, count = {}, {}, 0
    for (nc, h, w), (H, W) in shapes:
        x = torch.randn((1, nc, h, w), device=dev) * 3
        for ac in (False, True):
            ref = F.interpolate(x, size=(H, W), mode='bilinear', align_corners=ac)
            count += ref.numel()
            f32 = torch.float32
            sh = float(torch.tensor(h - 1 if ac else h, dtype=f32) / torch.tensor(H - 1 if ac else H, dtype=f32))
            sw = float(torch.tensor(w - 1 if ac else w, dtype=f32) / torch.tensor(W - 1 if ac else W, dtype=f32))
            out = torch.empty_like(ref)
            for form in (0, 1, 2):
                for idx_fma in (1, 0):
                    for outer in ((0, 1, 2) if form == 0 else ((0, 2) if form == 1 else (0,))):
                        for inner in ((0, 1, 2) if form == 0 else (0, 2)):
                            rc = lib.probe_launch(x.data_ptr(), out.data_ptr(), nc, h, w, H, W, int(ac), sh, sw, form, outer, inner,
                                                  idx_fma, torch.cuda.current_stream().cuda_stream)
                            assert rc == 0
                            torch.cuda.synchronize()
                            a, b = out.view(torch.int32).long(), ref.view(torch.int32).long()
                            bad = int((a != b).sum())
                            ulp = int((a - b).abs().max()) if bad else 0
                            key = (form, idx_fma, outer, inner)
                            tot[key] = tot.get(key, 0) + bad
                            worst[key] = max(worst.get(key, 0), ulp)
    print('elements compared per variant: %d' % count)
    for key in sorted(tot, key=lambda k: tot[k]):
        print('form=%d idx_fma=%d outer=%-14s inner=%-14s  mismatching: %9d  max ulp-ish diff: %d' % (
            key[0], key[1], names[key[2]], names[key[3]], tot[key], worst[key]))


if __name__ == '__main__':
    main()
