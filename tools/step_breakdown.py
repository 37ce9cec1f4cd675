#!/usr/bin/env python
"""Where the headline step's time goes beyond its dominant kernel: CUDA-graph replays of (a) forward only — memset,
up_gen (gradient sums included), finalize — and (b) forward + backward (adds autograd's fill and up_combine), over a
ring of input sets larger than L2.   python tools/step_breakdown.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import image_segmentation_lab_b200 as B  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    R = 8
    xs = [bench.make_logits((8, 19, 64, 128), 50 + i, device=dev).requires_grad_(True) for i in range(R)]
    ys = [bench.make_labels((8, 512, 1024), 19, 50 + i, 255, device=dev).unsqueeze(1) for i in range(R)]
    ce = B.CrossEntropyLoss()
    one = torch.ones((), device=dev)

    def step(i, mode):
        xs[i].grad = None
        loss = B.fused_resize_losses(xs[i], ys[i], ce, align_corners=False, ignore_index=255)['loss_ce']
        if mode == 'fwd_bwd':
            loss.backward()
        elif mode == 'fwd_bwd_given_grad':
            loss.backward(gradient=one)
        return loss

    for mode in ('fwd', 'fwd_bwd', 'fwd_bwd_given_grad'):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(R):
                step(i, mode)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs = []
        for i in range(R):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step(i, mode)
            graphs.append(g)
        for k in range(3 * R):
            graphs[k % R].replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 400
        s.record()
        for k in range(n):
            graphs[k % R].replay()
        e.record()
        torch.cuda.synchronize()
        print('%-20s %.2f us per step' % (mode, s.elapsed_time(e) / n * 1e3))


if __name__ == '__main__':
    main()
