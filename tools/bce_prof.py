"""Per-kernel device times and event-timed step of the sigmoid cross-entropy (python tools/bce_prof.py [N C H W])."""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

import image_segmentation_lab_b200 as B
from tests.helpers import synth_labels, synth_logits

N, C, H, W = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (32, 2, 512, 512)
x = synth_logits((N, C, H, W), 2, device='cuda', margin=False).requires_grad_(True)
y = synth_labels((N, H, W), C, 2, ignore_index=255, device='cuda')
mod = B.CrossEntropyLoss(use_sigmoid=True)


def step():
    x.grad = None
    mod(x, y, ignore_index=255).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    step()
e1.record()
torch.cuda.synchronize()
print('eager fwd+bwd %.1f us' % (e0.elapsed_time(e1) / 20 * 1e3))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
g.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    g.replay()
e1.record()
torch.cuda.synchronize()
print('graph fwd+bwd %.1f us' % (e0.elapsed_time(e1) / 20 * 1e3))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0]
for k, t, c in sorted(rows, key=lambda r: -r[1])[:12]:
    print('  %9.1f us total  x%-4d %s' % (t, c, k[:110]))
with torch.no_grad():
    for _ in range(3):
        mod(x, y, ignore_index=255)
    gf = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gf):
        mod(x, y, ignore_index=255)
    gf.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        gf.replay()
    e1.record()
    torch.cuda.synchronize()
    print('graph fwd only %.1f us' % (e0.elapsed_time(e1) / 20 * 1e3))
