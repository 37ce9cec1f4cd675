"""Times LovaszLoss forward(+backward) with CUDA events (python tools/ab_lovasz.py [N C H W dtype])."""
import sys
import time

import torch

import image_segmentation_lab_b200 as B
from tests.helpers import synth_labels, synth_logits


def run(N, C, H, W, dtype, per_image=False, iters=5):
    x = synth_logits((N, C, H, W), 2, dtype=dtype, device='cuda', margin=False).requires_grad_(True)
    y = synth_labels((N, H, W), C, 2, ignore_index=255, device='cuda')
    mod = B.LovaszLoss(per_image=per_image, reduction='mean' if per_image else 'none')
    res = {}
    for tag, grad in (('fwd', False), ('fwd_bwd', True)):
        def step():
            if grad:
                x.grad = None
                mod(x, y, ignore_index=255).backward()
            else:
                with torch.no_grad():
                    mod(x, y, ignore_index=255)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = B.launch_count()
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        res[tag] = e0.elapsed_time(e1) / iters
        res[tag + '_launches'] = (B.launch_count() - n0) // iters
    px = N * H * W
    print('lovasz N=%d C=%d %dx%d %s per_image=%s: fwd %.3f ms (%.1f Mpix/s), fwd+bwd %.3f ms (%.1f Mpix/s), own launches %d/%d'
          % (N, C, H, W, str(dtype).split('.')[-1], per_image, res['fwd'], px / res['fwd'] / 1e3, res['fwd_bwd'],
             px / res['fwd_bwd'] / 1e3, res['fwd_launches'], res['fwd_bwd_launches']), flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1:
        N, C, H, W = map(int, sys.argv[1:5])
        run(N, C, H, W, getattr(torch, sys.argv[5]) if len(sys.argv) > 5 else torch.float32)
    else:
        run(8, 19, 512, 1024, torch.float32)
        run(8, 19, 512, 1024, torch.float32, per_image=True)
        run(32, 21, 512, 512, torch.float32)
        run(16, 150, 512, 512, torch.bfloat16)
    if len(sys.argv) == 1:   # per-kernel device times of one forward+backward (CUPTI through torch.profiler)
        from torch.profiler import ProfilerActivity, profile
        x = synth_logits((8, 19, 512, 1024), 2, device='cuda', margin=False).requires_grad_(True)
        y = synth_labels((8, 512, 1024), 19, 2, ignore_index=255, device='cuda')
        mod = B.LovaszLoss(reduction='none')
        mod(x, y, ignore_index=255).backward()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            x.grad = None
            mod(x, y, ignore_index=255).backward()
            torch.cuda.synchronize()
        rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0]
        for k, t, c in sorted(rows, key=lambda r: -r[1])[:12]:
            print('  %9.1f us total  x%-4d %s' % (t, c, k[:120]))
