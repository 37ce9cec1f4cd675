#!/usr/bin/env python
"""Times BASELINE config 3 (bf16, 150 classes, 512x512, batch 16): class-weighted CE + Dice forward and forward+backward."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, image_segmentation_lab_b200 as B
dev=torch.device("cuda",0); peak,_=bench.hbm_peak()
def roof(a,ms): return a/(ms*1e-3)/1e9/peak
shape=(16,150,512,512)
xs=[bench.make_logits(shape,300+i,dtype=torch.bfloat16,device=dev).requires_grad_(True) for i in range(2)]
ys=[bench.make_labels((16,512,512),150,300+i,255,device=dev).unsqueeze(1) for i in range(2)]
cw=torch.linspace(0.5,1.5,150).tolist()
losses=[B.CrossEntropyLoss(class_weight=cw),B.DiceLoss(loss_weight=3.0)]
def fwd(i):
    with torch.no_grad(): B.fused_resize_losses(xs[i&1],ys[i&1],losses,ignore_index=255)
def fb(i):
    x=xs[i&1]; x.grad=None
    r=B.fused_resize_losses(x,ys[i&1],losses,ignore_index=255)
    (r["loss_ce"]+r["loss_dice"]).backward()
for i in range(3): fwd(i); fb(i)
a=bench.timed_events(fwd,10); b=bench.timed_events(fb,10)
el=16*150*512*512*2; px=16*512*512
print("C3 fwd %.3f ms (%.2f)  fwd+bwd %.3f ms (%.2f)"%(a,roof(el+px*8,a),b,roof(3*el+2*px*8,b)))
if os.environ.get("AB_KERNELS", "1") == "1":   # per-kernel device times (CUPTI through torch.profiler)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(4): fb(i)
        torch.cuda.synchronize()
    rows=[(e.key, e.device_time_total/max(e.count,1), e.count) for e in prof.key_averages() if e.device_time_total>0]
    for k,t,c in sorted(rows,key=lambda r:-r[1])[:8]: print("  %8.1f us x%d  %s"%(t,c,k[:110]))
