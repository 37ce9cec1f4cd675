"""CrossEntropyLoss with the reference's signatures (models/losses/cross_entropy_loss.py).

``cross_entropy`` :23-74 and ``CrossEntropyLoss`` :206-306 run on the fused sm_100a kernels
(csrc/loss_stream.cu, csrc/loss_bulk.cu, csrc/loss_rt.cuh, csrc/loss_upcell.cuh); ``binary_cross_entropy`` :100-164 (use_sigmoid) runs on
csrc/loss_bce.cu. ``mask_cross_entropy`` :167-203 is an instance-segmentation helper outside the
per-pixel path and is not provided.
"""
import warnings

import torch
import torch.nn as nn

from ._function import LossSpec, run_fused
from .utils import class_weight_tensor, get_class_weight


def _as_image(pred, label, weight):
    """(N,C) / (N,C,d1,...) predictions -> (N',C,H,W) plus a function restoring per-element shape."""
    if pred.dim() == 2:  # classification layout: one "image" whose pixels are the N samples
        n = pred.size(0)
        pred4 = pred.t().reshape(1, pred.size(1), 1, n)
        lab = label.reshape(1, 1, n)
        w = weight.reshape(1, 1, n) if weight is not None else None
        return pred4, lab, w, (lambda t: t.reshape(n))
    if pred.dim() == 3:
        pred4 = pred.unsqueeze(-1)
        lab = label.unsqueeze(-1)
        w = weight.unsqueeze(-1) if weight is not None else None
        return pred4, lab, w, (lambda t: t.squeeze(-1))
    if pred.dim() == 4:
        return pred, label, weight, (lambda t: t)
    n, c = pred.shape[:2]
    rest = tuple(pred.shape[2:])
    pred4 = pred.reshape(n, c, 1, -1)
    lab = label.reshape(n, 1, -1)
    w = weight.reshape(n, 1, -1) if weight is not None else None
    return pred4, lab, w, (lambda t: t.reshape((n,) + rest))


def cross_entropy(pred, label, weight=None, class_weight=None, reduction='mean', avg_factor=None,
                  ignore_index=-100, avg_non_ignore=False):
    """Softmax cross-entropy; same arguments and reduction rules as the reference (:23-74).

    * per-pixel loss ``-class_weight[y] * log_softmax(pred)[y]``, 0 where ``y == ignore_index`` (:56-61)
    * default 'mean' divides by ALL pixels; ``avg_non_ignore`` divides by the non-ignored count (:67-68)
    * ``weight`` is a per-pixel weight (cast to float, :69-70); ``avg_factor`` as in utils.py:72-79
    """
    if reduction not in ('none', 'mean', 'sum'):
        raise ValueError('%s is not a valid value for reduction' % reduction)
    if avg_factor is not None and reduction == 'sum':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    pred4, lab, w, restore = _as_image(pred, label, weight)
    spec = LossSpec(want_ce=True, ce_reduction=reduction,
                    ce_class_weight=class_weight_tensor(class_weight, pred.device),
                    ce_loss_weight=1.0, ce_ignore_index=int(ignore_index), ce_avg_non_ignore=bool(avg_non_ignore),
                    ce_avg_factor=None if avg_factor is None else float(avg_factor))
    loss, _, _ = run_fused(pred4, lab, w, spec)
    if reduction == 'none' or (avg_factor is not None and reduction == 'none'):
        loss = restore(loss)
    return _match_dtype(loss, pred)


def _match_dtype(loss, pred):
    """The reference returns the loss in pred's dtype, or float32 under autocast."""
    if pred.dtype != torch.float32 and not torch.is_autocast_enabled():
        return loss.to(pred.dtype)
    return loss


class CrossEntropyLoss(nn.Module):
    """Drop-in for the reference's ``CrossEntropyLoss`` (:206-306): same constructor, ``forward`` and
    ``loss_name``; owns no parameters or buffers (empty ``state_dict``)."""

    def __init__(self, use_sigmoid=False, use_mask=False, reduction='mean', class_weight=None, loss_weight=1.0,
                 loss_name='loss_ce', avg_non_ignore=False):
        super().__init__()
        assert (use_sigmoid is False) or (use_mask is False)
        self.use_sigmoid = use_sigmoid
        self.use_mask = use_mask
        self.reduction = reduction
        self.loss_weight = loss_weight
        self.class_weight = get_class_weight(class_weight)
        self.avg_non_ignore = avg_non_ignore
        if not self.avg_non_ignore and self.reduction == 'mean':
            warnings.warn(
                'Default ``avg_non_ignore`` is False, if you would like to ignore the certain label and average '
                'loss over non-ignore labels, which is the same with PyTorch official cross_entropy, set '
                '``avg_non_ignore=True``.')
        if self.use_mask:
            raise NotImplementedError('mask_cross_entropy (instance masks) is outside the per-pixel hot path')
        self._loss_name = loss_name
        self._cw_cache = {}
        self.single_pass = True

    def extra_repr(self):
        return f'avg_non_ignore={self.avg_non_ignore}'

    def _class_weight_on(self, device):
        if self.class_weight is None:
            return None
        t = self._cw_cache.get(device)
        if t is None:
            t = class_weight_tensor(self.class_weight, device)
            self._cw_cache[device] = t
        return t

    def spec(self, device, weight=None, avg_factor=None, reduction_override=None, ignore_index=-100):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if avg_factor is not None and reduction == 'sum':
            raise ValueError('avg_factor can not be used with reduction="sum"')
        return LossSpec(want_ce=True, ce_reduction=reduction, ce_class_weight=self._class_weight_on(device),
                        ce_loss_weight=float(self.loss_weight), ce_ignore_index=int(ignore_index),
                        ce_avg_non_ignore=bool(self.avg_non_ignore),
                        ce_avg_factor=None if avg_factor is None else float(avg_factor),
                        single_pass=self.single_pass)

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, ignore_index=-100,
                **kwargs):
        if self.use_sigmoid:
            from ._bce import binary_cross_entropy
            assert reduction_override in (None, 'none', 'mean', 'sum')
            reduction = reduction_override if reduction_override else self.reduction
            return binary_cross_entropy(
                cls_score, label, weight, class_weight=self._class_weight_on(cls_score.device), reduction=reduction,
                avg_factor=avg_factor, avg_non_ignore=self.avg_non_ignore, ignore_index=ignore_index,
                _loss_weight=float(self.loss_weight), _single_pass=bool(self.single_pass), **kwargs)
        spec = self.spec(cls_score.device, weight, avg_factor, reduction_override, ignore_index)
        pred4, lab, w, restore = _as_image(cls_score, label, weight)
        loss, _, _ = run_fused(pred4, lab, w, spec)
        if spec.ce_reduction == 'none':
            loss = restore(loss)
        return _match_dtype(loss, cls_score)

    @property
    def loss_name(self):
        """Key of this loss in the head's loss dict; names starting with ``loss_`` are back-propagated
        (utils/train_utils.py:53)."""
        return self._loss_name
