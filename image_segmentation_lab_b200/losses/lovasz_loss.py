"""LovaszLoss with the reference's signatures (models/losses/lovasz_loss.py:234-312) on csrc/loss_lovasz.cu.

``loss_type='multi_class'`` (Lovasz-Softmax, :135-231): the soft-max of :281-282 is fused in — the kernels read the raw
logits and the per-pixel log-sum-exp of one forward pass. For every class c to average ('present': classes with at least
one foreground pixel among the valid ones, 'all', or a list) over the valid pixels (label != ignore_index):
errors ``|1[y=c] - p_c|`` sorted descending, dotted with the first-differenced Jaccard index of the prefix sets
(``lovasz_grad`` :26-39); the class losses are weighted and averaged. ``loss_type='binary'`` (Lovasz hinge, :69-132) does
the same with errors ``1 - z * (2y - 1)`` through a relu. ``per_image=True`` evaluates each image separately and reduces
the per-image values with ``weight_reduce_loss`` (models/losses/utils.py:48-80).

Deliberate differences: the Jaccard increments are formed from exact integer counts (the reference's fp32
``J_i - J_{i-1}`` cancels; its fp32 cumsum is inexact above 2**24 pixels); a batch without a single valid pixel yields a
0-dim zero with zero gradient (the reference returns an EMPTY tensor, ``probs * 0.`` :148-150); with ``'present'`` and
valid pixels but no in-range label the reference raises from ``torch.stack([])`` — here the loss is 0.
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib
from ._function import prep_labels
from .cross_entropy_loss import _match_dtype
from .utils import class_weight_tensor, get_class_weight


class _LovaszFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, cfg, grad_enabled=True):
        lib = _lib.load()
        _lib.require_cuda(logits, "cls_score")
        if logits.dtype not in _lib.LOGIT_DTYPES:
            raise TypeError("cls_score must be float32, bfloat16 or float16, got %s" % logits.dtype)
        binary = cfg['binary']
        x = logits.contiguous()
        labels = prep_labels(labels, x) if x.dim() == 4 else labels.to(x.device).contiguous()
        if labels.dtype not in _lib.LABEL_DTYPES:
            labels = labels.long()
        N = int(x.shape[0])
        if binary:
            Cc = 1
            HW = x.numel() // N if N else 0
            if labels.numel() != x.numel():
                raise ValueError("binary Lovasz hinge: logits %s and labels %s must have the same number of elements"
                                 % (tuple(x.shape), tuple(labels.shape)))
        else:
            if x.dim() != 4:
                raise ValueError("cls_score must be (N,C,H,W), got %s" % (tuple(x.shape),))
            Cc = int(x.shape[1])
            HW = int(x.shape[2] * x.shape[3])
            if labels.numel() != N * HW:
                raise ValueError("label %s does not match cls_score %s" % (tuple(labels.shape), tuple(x.shape)))
        dev = x.device
        needs_grad = bool(ctx.needs_input_grad[0]) and bool(grad_enabled)
        per_image = bool(cfg['per_image'])
        n_groups = N if per_image else 1
        n_seg = 1 if binary else Cc
        red = cfg['reduction']
        none_vec = per_image and red == 'none'
        stream = _lib.stream_ptr(dev)
        keep = []
        with torch.cuda.device(dev):
            lse = None
            if not binary and N * HW > 0:
                # one pass over the logits: per-pixel log-sum-exp (the CE statistics it also produces are not used)
                lse = torch.empty((N, HW), dtype=torch.float32, device=dev)
                st = torch.empty(_lib.STATS_WORDS, dtype=torch.int64, device=dev)
                fd = _lib.LossDesc()
                fd.logits = x.data_ptr(); fd.labels = labels.data_ptr()
                fd.logit_dtype = _lib.LOGIT_DTYPES[x.dtype]; fd.label_dtype = _lib.LABEL_DTYPES[labels.dtype]
                fd.N, fd.C, fd.h, fd.w, fd.H, fd.W = N, Cc, int(x.shape[2]), int(x.shape[3]), int(x.shape[2]), int(x.shape[3])
                fd.flags = _lib.WANT_CE | _lib.WANT_LSE
                fd.ignore_index = -100
                fd.lse = lse.data_ptr(); fd.stats = st.data_ptr()
                fd.ce_loss_weight = 1.0
                _lib.check(lib.b200seg_loss_fwd(C.byref(fd), stream))
                keep.append(st)
            pairs = int(binary or needs_grad)
            ws_bytes = int(lib.b200seg_lovasz_workspace_bytes(N, n_seg, HW, int(per_image), pairs)) if N * HW > 0 else 256
            if ws_bytes < 0:
                raise RuntimeError(_lib.last_error())
            ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
            ws_ptr = (ws.data_ptr() + 255) & ~255
            lab16 = torch.empty((N, HW), dtype=torch.int16, device=dev)
            G = None
            if needs_grad:
                G = torch.empty((N, n_seg, HW), dtype=torch.float32, device=dev)
            small = torch.empty(max(n_groups, 1) * n_seg * 2 + max(n_groups, 1) * (n_seg + 1) // 2 + 2, dtype=torch.float64, device=dev)
            seg_stats = small[:max(n_groups, 1) * n_seg * 2]
            f32 = small[max(n_groups, 1) * n_seg * 2:].view(torch.float32)
            coef = f32[:max(n_groups, 1) * n_seg]
            # the result is a tensor of its own (not a view of `small`): `loss[name] += ...` in the stock head
            # (decode_head.py:290) updates it in place, which autograd refuses for a view made inside a Function
            out = torch.empty(() if not none_vec else (n_groups,), dtype=torch.float32, device=dev)

            d = _lib.LovaszDesc()
            d.logits = x.data_ptr(); d.labels = labels.data_ptr()
            d.lse = lse.data_ptr() if lse is not None else None
            cw = cfg['class_weight']
            d.class_weight = cw.data_ptr() if (cw is not None and not binary) else None
            d.logit_dtype = _lib.LOGIT_DTYPES[x.dtype]; d.label_dtype = _lib.LABEL_DTYPES[labels.dtype]
            d.N, d.C, d.HW = N, Cc, HW
            ign = cfg['ignore_index']
            d.has_ignore = int(ign is not None)
            d.ignore_index = int(ign) if ign is not None else 0
            d.binary = int(binary); d.per_image = int(per_image)
            classes = cfg['classes']
            d.only_present = int(classes == 'present')
            if isinstance(classes, (list, tuple)) and not binary:
                arr = (C.c_int32 * len(classes))(*[int(c) for c in classes])
                keep.append(arr)
                d.classes_host = C.cast(arr, C.POINTER(C.c_int32))
                d.n_classes = len(classes)
            d.reduction = _lib.REDUCTIONS[red]
            d.has_avg_factor = int(cfg['avg_factor'] is not None and per_image)
            d.avg_factor = float(cfg['avg_factor'] or 0.0)
            d.loss_weight = float(cfg['loss_weight'])
            d.lab16 = lab16.data_ptr()
            d.G = G.data_ptr() if G is not None else None
            d.workspace = ws_ptr; d.workspace_bytes = ws_bytes
            d.seg_stats = seg_stats.data_ptr(); d.out = out.data_ptr()
            d.coef = coef.data_ptr() if needs_grad else None
            _lib.check(lib.b200seg_lovasz_fwd(C.byref(d), stream))
        result = out
        if needs_grad:
            ctx.cfg = dict(binary=binary, per_image=per_image, none_vec=none_vec, N=N, C=Cc, HW=HW)
            ctx.small = small
            ctx.save_for_backward(x, lse if lse is not None else small, lab16, G, coef)
        return result

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        x, lse, lab16, G, coef = ctx.saved_tensors
        cfg = ctx.cfg
        dev = x.device
        out = torch.empty_like(x)
        if g_out is None:
            return out.zero_(), None, None, None
        with torch.cuda.device(dev):
            go = g_out.detach().to(torch.float32).contiguous()
            b = _lib.LovaszBwdDesc()
            b.logits = x.data_ptr()
            b.lse = None if cfg['binary'] else lse.data_ptr()
            b.lab16 = lab16.data_ptr(); b.G = G.data_ptr(); b.coef = coef.data_ptr()
            b.grad_out = go.data_ptr()
            b.grad_logits = out.data_ptr()
            b.logit_dtype = _lib.LOGIT_DTYPES[x.dtype]
            b.N, b.C, b.HW = cfg['N'], cfg['C'], cfg['HW']
            b.binary = int(cfg['binary']); b.per_image = int(cfg['per_image'])
            b.grad_per_group = int(cfg['none_vec'])
            _lib.check(lib.b200seg_lovasz_bwd(C.byref(b), _lib.stream_ptr(dev)))
        return out, None, None, None


def _is_list_of_int(seq):
    return isinstance(seq, list) and all(isinstance(c, int) for c in seq)


class LovaszLoss(nn.Module):
    """Drop-in for the reference's ``LovaszLoss`` (:234-312): same constructor, ``forward`` and ``loss_name``."""

    def __init__(self, loss_type='multi_class', classes='present', per_image=False, reduction='mean', class_weight=None,
                 loss_weight=1.0, loss_name='loss_lovasz'):
        super().__init__()
        assert loss_type in ('binary', 'multi_class'), "loss_type should be 'binary' or 'multi_class'."   # :263-264
        assert classes in ('all', 'present') or _is_list_of_int(classes)                                  # :270
        if not per_image:
            assert reduction == 'none', "reduction should be 'none' when per_image is False."           # :271-273
        if _is_list_of_int(classes) and len(set(classes)) != len(classes):
            raise ValueError('classes lists a class more than once: %r' % (classes,))
        self.loss_type = loss_type
        self.classes = classes
        self.per_image = per_image
        self.reduction = reduction
        self.loss_weight = loss_weight
        self.class_weight = get_class_weight(class_weight)
        self._loss_name = loss_name
        self._cw_cache = {}

    def _class_weight_on(self, device):
        if self.class_weight is None:
            return None
        t = self._cw_cache.get(device)
        if t is None:
            t = class_weight_tensor(self.class_weight, device)
            self._cw_cache[device] = t
        return t

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, ignore_index=255, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')                                       # :284
        reduction = reduction_override if reduction_override else self.reduction
        if self.per_image and avg_factor is not None and reduction == 'sum':
            raise ValueError('avg_factor can not be used with reduction="sum"')                        # utils.py:78-79
        binary = self.loss_type == 'binary'
        if binary and cls_score.dim() == 4 and cls_score.size(1) != 1:
            raise ValueError('Sigmoid output possible only with 1 class')                               # :158
        cfg = dict(binary=binary, classes=self.classes, per_image=self.per_image, reduction=reduction,
                   class_weight=self._class_weight_on(cls_score.device), loss_weight=float(self.loss_weight),
                   avg_factor=avg_factor, ignore_index=ignore_index)
        loss = _LovaszFunction.apply(cls_score, label, cfg, torch.is_grad_enabled())
        return _match_dtype(loss, cls_score)

    @property
    def loss_name(self):
        return self._loss_name
