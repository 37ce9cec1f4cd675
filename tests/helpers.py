"""Shared adaptors: run one golden / synthetic case through the ORACLE or through the CUDA package."""
import warnings

import numpy as np
import torch

from oracle import oracle as O


def case_tensors(data, name, device='cpu', dtype=torch.float32):
    logits = torch.from_numpy(data[name + '/logits']).to(device=device, dtype=dtype)
    labels = torch.from_numpy(data[name + '/labels']).to(device)
    pw = torch.from_numpy(data[name + '/pixel_weight']).to(device) if (name + '/pixel_weight') in data else None
    return logits, labels, pw


def loss_case_oracle(case, logits, labels, pw, grad_out=None):
    """Returns dict(loss.. , grad, acc) computed by the oracle on logits' device."""
    warnings.simplefilter('ignore')
    logits = logits.detach().clone().requires_grad_(True)
    H, W = case['size']
    ac = case.get('ac', False)
    full = O.resize(logits, size=(H, W), mode='bilinear', align_corners=ac)
    out = {}
    if case['kind'] == 'ce':
        loss = O.cross_entropy_loss_module(full, labels, weight=pw, avg_factor=case.get('avg_factor'),
                                           ignore_index=case['ignore'], **case['kw'])
        out['loss'] = loss
        total = (loss * grad_out).sum() if grad_out is not None else (loss.sum() if loss.dim() else loss)
    elif case['kind'] == 'dice':
        loss = O.dice_loss_module(full, labels, avg_factor=case.get('avg_factor'), **case['kw'])
        out['loss'] = loss
        total = loss
    else:
        l1 = O.cross_entropy_loss_module(full, labels, weight=pw, ignore_index=case['ignore'], **case['ce'])
        l2 = O.dice_loss_module(full, labels, **case['dice'])
        out['loss_ce'], out['loss_dice'] = l1, l2
        total = l1 + l2
    total.backward()
    out['grad'] = logits.grad
    out['acc'] = O.accuracy(full.detach(), labels, ignore_index=case['ignore'] if case['ignore'] != -100 else None)
    return {k: v.detach() for k, v in out.items()}


def loss_case_cuda(case, logits, labels, pw, grad_out=None, label_dtype=None, single_pass=True):
    """Same quantities through image_segmentation_lab_b200 (module API, resize fused via fused_resize_losses)."""
    import image_segmentation_lab_b200 as B
    warnings.simplefilter('ignore')
    logits = logits.detach().clone().requires_grad_(True)
    if label_dtype is not None:
        labels = labels.to(label_dtype)
    ac = case.get('ac', False)
    ign = case['ignore']
    out = {}
    seg_label = labels.unsqueeze(1)
    if case['kind'] == 'ce':
        mod = B.CrossEntropyLoss(**case['kw'])
        mod.single_pass = single_pass
        if case.get('avg_factor') is not None or mod.reduction == 'none':
            full = B.resize(logits, size=case['size'], mode='bilinear', align_corners=ac, warning=False) \
                if tuple(logits.shape[2:]) != tuple(case['size']) else logits
            loss = mod(full, labels, weight=pw, avg_factor=case.get('avg_factor'), ignore_index=ign)
            acc = B.accuracy(full.detach(), labels, ignore_index=ign if ign != -100 else None)
        else:
            r = B.fused_resize_losses(logits, seg_label, mod, align_corners=ac, ignore_index=ign, seg_weight=pw)
            loss, acc = r[mod.loss_name], r['acc_seg']
            if ign == -100:
                full = B.resize(logits.detach(), size=case['size'], mode='bilinear', align_corners=ac, warning=False)
                acc = B.accuracy(full, labels, ignore_index=None)
        out['loss'] = loss
        total = (loss * grad_out).sum() if grad_out is not None else (loss.sum() if loss.dim() else loss)
    elif case['kind'] == 'dice':
        mod = B.DiceLoss(**case['kw'])
        if case.get('avg_factor') is not None:
            loss = mod(logits, labels, avg_factor=case.get('avg_factor'))
            acc = B.accuracy(logits.detach(), labels, ignore_index=ign)
        else:
            r = B.fused_resize_losses(logits, seg_label, mod, align_corners=ac, ignore_index=ign)
            loss, acc = r[mod.loss_name], r['acc_seg']
        out['loss'] = loss
        total = loss
    else:
        ce = B.CrossEntropyLoss(**case['ce'])
        ce.single_pass = single_pass
        dice = B.DiceLoss(**case['dice'])
        r = B.fused_resize_losses(logits, seg_label, [ce, dice], align_corners=ac, ignore_index=ign, seg_weight=pw)
        out['loss_ce'], out['loss_dice'] = r['loss_ce'], r['loss_dice']
        acc = r['acc_seg']
        total = r['loss_ce'] + r['loss_dice']
    total.backward()
    out['grad'] = logits.grad
    out['acc'] = acc
    return {k: v.detach() for k, v in out.items()}


def rel_err(a, b):
    """||a - b||_inf / max(||b||_inf, tiny) — the gate of BASELINE.md section 5."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    denom = max(float(b.abs().max()) if b.numel() else 0.0, 1e-30)
    return float((a - b).abs().max()) / denom if a.numel() else 0.0


def synth_logits(shape, seed, dtype=torch.float32, device='cpu', margin=True):
    """SURVEY.md 8d generator: randn*2, +1 on one class per pixel, quantised to 2**-6 (no soft-max rounding ties),
    plus a per-class offset of c * 2**-10 so that no two classes of a pixel are exactly equal."""
    g = torch.Generator().manual_seed(seed)
    n, c, h, w = shape
    x = torch.randn(shape, generator=g) * 2.0
    if margin:
        hot = torch.randint(0, c, (n, 1, h, w), generator=g)
        x.scatter_add_(1, hot, torch.ones((n, 1, h, w)))
        x = torch.round(x * 64.0) / 64.0
        x = x + (torch.arange(c, dtype=torch.float32).view(1, c, 1, 1) * (2.0 ** -10 if dtype == torch.float32 else 0.0))
    return x.to(dtype=dtype, device=device)


def synth_labels(shape, num_classes, seed, ignore_index=255, ignore_frac=0.1, block=16, dtype=torch.int64, device='cpu'):
    """randint labels in block x block constant tiles with ~ignore_frac of the tiles set to ignore_index."""
    g = torch.Generator().manual_seed(seed + 7)
    n, h, w = shape
    bh, bw = (h + block - 1) // block, (w + block - 1) // block
    y = torch.randint(0, num_classes, (n, bh, bw), generator=g)
    if ignore_index is not None and ignore_frac > 0:
        y[torch.rand((n, bh, bw), generator=g) < ignore_frac] = ignore_index
    y = y.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :h, :w].contiguous()
    return y.to(dtype=dtype, device=device)
