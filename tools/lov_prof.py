"""Per-kernel device times of one LovaszLoss forward+backward (python tools/lov_prof.py [N C H W dtype])."""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

import image_segmentation_lab_b200 as B
from tests.helpers import synth_labels, synth_logits

N, C, H, W = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (8, 19, 512, 1024)
dtype = getattr(torch, sys.argv[5]) if len(sys.argv) > 5 else torch.float32
x = synth_logits((N, C, H, W), 2, dtype=dtype, device='cuda', margin=False).requires_grad_(True)
y = synth_labels((N, H, W), C, 2, ignore_index=255, device='cuda')
mod = B.LovaszLoss(reduction='none')
for _ in range(2):
    x.grad = None
    mod(x, y, ignore_index=255).backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    x.grad = None
    mod(x, y, ignore_index=255).backward()
e1.record()
torch.cuda.synchronize()
print('fwd+bwd %.3f ms' % (e0.elapsed_time(e1) / 5))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    x.grad = None
    mod(x, y, ignore_index=255).backward()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0]
for k, t, c in sorted(rows, key=lambda r: -r[1])[:8]:
    print('  %9.1f us total  x%-4d %s' % (t, c, k[:110]))
