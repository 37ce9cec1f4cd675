#!/usr/bin/env python
"""Runs tools/probe/fma_probe.cu against F.interpolate(mode='bilinear') on this GPU: prints, per contraction variant, how
many output elements differ bitwise from ATen over a set of non-power-of-two shapes (both align_corners settings)."""
import ctypes
import os
import subprocess
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'libfma_probe.so')


def build():
    src = os.path.join(HERE, 'fma_probe.cu')
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-shared', '-Xcompiler', '-fPIC', '-cudart',
                        'static', src, '-o', SO], check=True)
    return ctypes.CDLL(SO)


def main():
    lib = build()
    if not torch.cuda.is_available():
        print('built', SO)
        return
    dev = torch.device('cuda', 0)
    shapes = [((3, 19, 47), (513, 1025)), ((2, 65, 129), (513, 1025)), ((4, 64, 128), (512, 1024)), ((5, 33, 31), (100, 77)),
              ((2, 128, 256), (1024, 2048)), ((7, 60, 90), (61, 91)), ((3, 100, 100), (37, 53))]
    names = {0: 'fma(x,a,y*b)', 1: 'fma(y,b,x*a)', 2: 'no fma'}
    tot = {}
    for (nc, h, w), (H, W) in shapes:
        x = torch.randn((1, nc, h, w), device=dev) * 3
        for ac in (False, True):
            ref = F.interpolate(x, size=(H, W), mode='bilinear', align_corners=ac)
            if ac:
                sh = (h - 1) / (H - 1) if H > 1 else 0.0
                sw = (w - 1) / (W - 1) if W > 1 else 0.0
            else:
                sh, sw = h / H, w / W
            sh = float(torch.tensor(h - 1 if ac else h, dtype=torch.float32) / torch.tensor(H - 1 if ac else H, dtype=torch.float32))
            sw = float(torch.tensor(w - 1 if ac else w, dtype=torch.float32) / torch.tensor(W - 1 if ac else W, dtype=torch.float32))
            out = torch.empty_like(ref)
            for idx_fma in (1, 0):
                for outer in (0, 1, 2):
                    for inner in (0, 1, 2):
                        rc = lib.probe_launch(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), nc, h, w, H, W, int(ac),
                                              ctypes.c_float(sh), ctypes.c_float(sw), outer, inner, idx_fma,
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
                        assert rc == 0
                        torch.cuda.synchronize()
                        bad = int((out.view(torch.int32) != ref.view(torch.int32)).sum())
                        key = (idx_fma, outer, inner)
                        tot[key] = tot.get(key, 0) + bad
    for key in sorted(tot, key=lambda k: tot[k]):
        print('idx_fma=%d outer=%-14s inner=%-14s  mismatching elements: %d' % (key[0], names[key[1]], names[key[2]], tot[key]))


if __name__ == '__main__':
    main()
