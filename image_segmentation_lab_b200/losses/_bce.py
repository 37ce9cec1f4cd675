"""Sigmoid cross-entropy with the reference's signature (models/losses/cross_entropy_loss.py:77-164) on
csrc/loss_bce.cu: the one-hot target, the valid mask and the expanded pixel weight are formed inside the kernel."""
import ctypes as C

import torch

from .. import _lib
from ._function import prep_labels

_EPS = float(torch.finfo(torch.float32).eps)


class _BceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, label, weight, pos_weight, reduction, avg_factor, ignore_index, avg_non_ignore, single,
                grad_enabled=True, loss_weight=1.0, single_pass=True, acc_out=None, acc_ignore=None):
        lib = _lib.load()
        _lib.require_cuda(pred, 'pred')
        if pred.dtype not in _lib.LOGIT_DTYPES:
            raise TypeError('pred must be float32, bfloat16 or float16, got %s' % pred.dtype)
        x = pred.contiguous()
        N, Cc = x.shape[0], x.shape[1]
        HW = x[0, 0].numel()
        dev = x.device
        lab = label
        if lab.dtype not in _lib.LABEL_DTYPES:
            lab = lab.long()
        lab = lab.to(dev).contiguous()
        assert lab.numel() == N * HW, 'label must have one entry per pixel'
        w = None
        if weight is not None:
            w = weight.to(device=dev, dtype=torch.float32).contiguous()
            assert w.numel() == N * HW, 'weight must have one entry per pixel'
        needs_grad = bool(ctx.needs_input_grad[0]) and bool(grad_enabled)
        use_nvalid = bool(reduction == 'mean' and avg_factor is None and avg_non_ignore)
        n_elem = max(N * Cc * HW, 1)
        scale = 1.0
        if reduction == 'mean':
            if avg_factor is not None:
                scale = 1.0 / float(torch.tensor(avg_factor + _EPS, dtype=torch.float32))
            elif not use_nvalid:
                scale = 1.0 / n_elem
        scale *= float(loss_weight)
        # single pass (the gradient is written by the forward kernel, with the upstream gradient taken as 1) whenever the
        # denominator is known before the launch; never for float16: a gradient formed at ~1/numel would be flushed
        # before a GradScaler's factor arrives (same rule as the soft-max path's flat plan)
        fused = bool(needs_grad and single_pass and reduction != 'none' and not use_nvalid and x.dtype != torch.float16
                     and N * HW > 0)
        with torch.cuda.device(dev):
            stats = torch.empty(8, dtype=torch.int64, device=dev)
            loss_elem = torch.empty(x.shape, dtype=torch.float32, device=dev) if reduction == 'none' else None
            out = None if reduction == 'none' else torch.empty((), dtype=torch.float32, device=dev)
            grad = torch.empty_like(x) if fused else None
            d = _lib.BceDesc()
            d.logits = x.data_ptr(); d.labels = lab.data_ptr()
            d.pixel_weight = w.data_ptr() if w is not None else None
            d.pos_weight = pos_weight.data_ptr() if pos_weight is not None else None
            d.logit_dtype = _lib.LOGIT_DTYPES[x.dtype]; d.label_dtype = _lib.LABEL_DTYPES[lab.dtype]
            d.N, d.C, d.HW = N, Cc, HW
            d.ignore_index = int(ignore_index)
            d.single_channel = int(bool(single))
            d.use_nvalid = int(use_nvalid)
            d.loss_weight = float(loss_weight)
            d.loss_elem = loss_elem.data_ptr() if loss_elem is not None else None
            d.stats = stats.data_ptr()
            d.out = out.data_ptr() if out is not None else None
            d.out_scale_host = float(scale)
            if acc_out is not None:      # top-1 accuracy of the same launch (decode_head.py:295)
                assert acc_out.dtype == torch.float32 and acc_out.is_cuda and acc_out.numel() == 1
                d.acc_out = acc_out.data_ptr()
                d.acc_has_ignore = int(acc_ignore is not None)
                d.acc_ignore_index = int(acc_ignore) if acc_ignore is not None else 0
            if fused:
                d.grad_logits = grad.data_ptr()
                d.grad_scale_host = float(scale)
            _lib.check(lib.b200seg_bce_fwd(C.byref(d), _lib.stream_ptr(dev)))
            if reduction == 'none':
                out = loss_elem
        if needs_grad:
            ctx.save_for_backward(x, lab, w if w is not None else stats, stats, grad if fused else stats)
            ctx.has_w = w is not None
            ctx.pos_weight = pos_weight
            ctx.fused = fused
            ctx.consumed = False
            ctx.cfg = (reduction, scale, use_nvalid, int(ignore_index), bool(single))
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, lab, w, stats, fgrad = ctx.saved_tensors
        reduction, scale, use_nvalid, ignore_index, single = ctx.cfg
        N, Cc = x.shape[0], x.shape[1]
        HW = x[0, 0].numel()
        dev = x.device
        none = (None,) * 13
        with torch.cuda.device(dev):
            if ctx.fused:
                if ctx.consumed:
                    raise RuntimeError("the single-pass loss graph can be back-propagated once; build the loss with "
                                       "single_pass=False to call backward() repeatedly (retain_graph)")
                ctx.consumed = True
                gs = g.detach().to(torch.float32).reshape(()).contiguous()
                _lib.check(lib.b200seg_scale_inplace(fgrad.data_ptr(), _lib.LOGIT_DTYPES[fgrad.dtype], fgrad.numel(),
                                                     gs.data_ptr(), _lib.stream_ptr(dev)))
                return (fgrad,) + none
            grad = torch.empty_like(x)
            d = _lib.BceDesc()
            d.logits = x.data_ptr(); d.labels = lab.data_ptr()
            d.pixel_weight = w.data_ptr() if ctx.has_w else None
            d.pos_weight = ctx.pos_weight.data_ptr() if ctx.pos_weight is not None else None
            d.logit_dtype = _lib.LOGIT_DTYPES[x.dtype]; d.label_dtype = _lib.LABEL_DTYPES[lab.dtype]
            d.N, d.C, d.HW = N, Cc, HW
            d.ignore_index = ignore_index
            d.single_channel = int(single)
            d.use_nvalid = int(use_nvalid)
            d.grad_scale_host = float(scale)
            keep = g.detach().to(torch.float32)
            if reduction == 'none':
                keep = keep.expand(x.shape).contiguous()
                d.grad_elem = keep.data_ptr()
            else:
                keep = keep.reshape(()).contiguous()
                d.grad_out = keep.data_ptr()
            d.grad_logits = grad.data_ptr()
            d.stats = stats.data_ptr()
            _lib.check(lib.b200seg_bce_bwd(C.byref(d), _lib.stream_ptr(dev)))
        return (grad,) + none


def binary_cross_entropy(pred, label, weight=None, reduction='mean', avg_factor=None, class_weight=None,
                         ignore_index=-100, avg_non_ignore=False, _loss_weight=1.0, _single_pass=True, _acc_out=None,
                         _acc_ignore=None, **kwargs):
    """Same arguments and results as the reference (:100-164). ``pred`` (N,C,H,W) with ``label`` (N,H,W) expands
    the label to one-hot inside the kernel; ``pred`` (N,1,H,W) treats the label (0/1) as the target (:126-134);
    ``pred`` and ``label`` of equal shape use the label as a soft target mask as the reference does.
    ``_loss_weight`` / ``_single_pass`` are this package's own (CrossEntropyLoss folds its loss_weight into the launch);
    ``_acc_out`` (a (1,) float32 CUDA tensor) receives the top-1 accuracy of ``pred`` against ``label`` over the pixels
    whose label is not ``_acc_ignore`` from the same launch (fused_resize_losses: decode_head.py:295)."""
    if reduction not in ('none', 'mean', 'sum'):
        raise ValueError('%s is not a valid value for reduction' % reduction)
    if avg_factor is not None and reduction == 'sum':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    single = False
    if pred.size(1) == 1 and pred.dim() == label.dim() + 1:
        single = True                       # the reference squeezes the channel (:133) and checks label <= 1 (:130)
    elif pred.dim() == label.dim():
        # element-wise targets: treat every element as its own pixel of a 1-channel prediction. The reference forwards
        # pos_weight=class_weight here too (:160-161); a scalar weight maps onto the single channel, anything else would
        # broadcast against the LAST dimension of pred, which the kernel's per-class table cannot express.
        shape = pred.shape
        pos_w = None
        if class_weight is not None:
            pos_w = torch.as_tensor(class_weight, dtype=torch.float32, device=pred.device).reshape(-1).contiguous()
            if pos_w.numel() != 1:
                raise NotImplementedError('binary_cross_entropy with element-wise targets supports a scalar class_weight only '
                                          '(got %d entries)' % pos_w.numel())
        out = _BceFunction.apply(pred.reshape(-1, 1, 1), label.reshape(-1, 1), None if weight is None else weight.reshape(-1, 1),
                                 pos_w, reduction, avg_factor, ignore_index, avg_non_ignore, True, torch.is_grad_enabled(),
                                 float(_loss_weight), bool(_single_pass))
        if reduction == 'none':
            out = out.reshape(shape)
        if pred.dtype != torch.float32 and not torch.is_autocast_enabled():
            out = out.to(pred.dtype)
        return out
    else:
        assert (pred.dim() == 2 and label.dim() == 1) or (pred.dim() == 4 and label.dim() == 3), \
            'Only pred shape [N, C], label shape [N] or pred shape [N, C, H, W], label shape [N, H, W] are supported'
    x = pred if pred.dim() >= 3 else pred.unsqueeze(-1)
    pos_w = None
    if class_weight is not None:
        pos_w = torch.as_tensor(class_weight, dtype=torch.float32, device=pred.device).contiguous()
    out = _BceFunction.apply(x, label, weight, pos_w, reduction, avg_factor, ignore_index, avg_non_ignore, single,
                             torch.is_grad_enabled(), float(_loss_weight), bool(_single_pass), _acc_out, _acc_ignore)
    if reduction == 'none':
        out = out.reshape(pred.shape)
        if single:
            out = out.squeeze(1)
    if pred.dtype != torch.float32 and not torch.is_autocast_enabled():
        out = out.to(pred.dtype)
    return out
