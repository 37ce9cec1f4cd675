"""GPU box: times the ORACLE (the reference's unfused ATen chain, oracle/oracle.py) on CUDA tensors for the BASELINE
configs — "the kernel to beat on the same box" of SURVEY.md 8d. Test infrastructure (lives under tests/ because only
tests/, smoke() and bench.py's baseline leg may execute oracle/); not collected by pytest.

    python tests/aten_same_gpu.py
"""
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402


def timed(fn, iters):
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    warnings.simplefilter('ignore')
    dev = torch.device('cuda', 0)
    rows = []

    def loss_case(name, shape, size, dtype, losses, iters):
        n = shape[0]
        x = bench.make_logits(shape, 7, dtype=dtype, device=dev).requires_grad_(True)
        y = bench.make_labels((n,) + size, shape[1], 7, device=dev).unsqueeze(1)

        def step():
            x.grad = None
            out = O.head_losses(x, y, losses, align_corners=False, ignore_index=255)
            tot = None
            for k, v in out.items():
                if k.startswith('loss'):
                    tot = v if tot is None else tot + v
            tot.backward()

        ms = timed(step, iters)
        rows.append((name, ms, n * size[0] * size[1] / ms / 1e3))
        del x, y
        torch.cuda.empty_cache()

    cw = torch.linspace(0.5, 1.5, 150).tolist()
    loss_case('C2 resize+CE fwd+bwd (8,19,64,128)->512x1024 fp32', (8, 19, 64, 128), (512, 1024), torch.float32,
              [('ce', {}, 'loss_ce')], 10)
    loss_case('C3 CE+Dice fwd+bwd (16,150,512,512) bf16', (16, 150, 512, 512), (512, 512), torch.bfloat16,
              [('ce', dict(class_weight=cw), 'loss_ce'), ('dice', dict(loss_weight=3.0), 'loss_dice')], 3)
    loss_case('C4 CE fwd+bwd (32,21,512,512) fp32', (32, 21, 512, 512), (512, 512), torch.float32, [('ce', {}, 'loss_ce')], 10)

    x = bench.make_logits((8, 19, 512, 1024), 9, device=dev).requires_grad_(True)
    y = bench.make_labels((8, 512, 1024), 19, 9, device=dev)

    def lov():
        x.grad = None
        O.lovasz_loss_module(x, y, reduction='none', ignore_index=255).backward()

    ms = timed(lov, 3)
    rows.append(('Lovasz-Softmax fwd+bwd (8,19,512,1024) fp32', ms, 8 * 512 * 1024 / ms / 1e3))
    del x, y
    torch.cuda.empty_cache()

    # C5 (i): intersect_and_union on 20 of the 500 label maps (the reference moves every result to the host per image)
    preds = [torch.randint(0, 19, (1024, 2048), device=dev) for _ in range(20)]
    gts = [bench.make_labels((1, 1024, 2048), 19, 50 + i, device=dev)[0].float() for i in range(20)]

    def iau():
        O.intersect_and_union_torch(preds, gts, 19, 255)

    ms = timed(iau, 3)
    rows.append(('C5(i) intersect_and_union, 20 x 1024x2048 label maps', ms, 20 * 1024 * 2048 / ms / 1e3))
    for name, ms, mp in rows:
        print('%-62s %10.3f ms  %12.1f Mpix/s' % (name, ms, mp), flush=True)


if __name__ == '__main__':
    main()
