"""DiceLoss with the reference's signatures (models/losses/dice_loss.py:23-148), on csrc/loss_rt.cuh (C <= 32) and csrc/loss_dice.cu (C > 32).

Quirks of the reference that are preserved:
  * the denominator sum(p^e + t^e) is NOT masked by valid_mask (:56);
  * labels are clamped to [0, C-1] for the one-hot, so ignored pixels (255) count in class C-1's
    sum(t) while being masked out of the numerator (:119-122);
  * class i == ignore_index is skipped but the divisor stays num_classes (:35,:45);
  * per class, the per-sample dice terms are averaged over the batch (binary_dice_loss is itself
    @weighted_loss with the default 'mean', :48);
  * ``weight=`` / ``ignore_index=`` passed at the call site are swallowed by **kwargs (:103-108).
"""
import torch
import torch.nn as nn

from ._function import LossSpec, run_fused
from .utils import class_weight_tensor, get_class_weight
from .cross_entropy_loss import _match_dtype


def dice_loss(pred, target, valid_mask=None, weight=None, reduction='mean', avg_factor=None, smooth=1, exponent=2,
              class_weight=None, ignore_index=255):
    """Functional form on raw logits ``pred`` (N,C,H,W) and integer labels ``target`` (N,H,W).

    Unlike the reference helper (:23-45), which receives soft-max probabilities, a one-hot target and a
    valid mask built by ``DiceLoss.forward``, this takes what ``DiceLoss.forward`` takes and computes the
    soft-max, one-hot and mask inside the kernel; ``valid_mask`` must be None (it is ``target != ignore_index``).
    """
    if valid_mask is not None:
        raise ValueError('valid_mask is derived from ignore_index inside the kernel; pass None')
    if weight is not None:
        raise ValueError('element-wise weight on the (scalar) dice loss is not supported')
    if avg_factor is not None and reduction == 'sum':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    spec = LossSpec(want_dice=True, dice_reduction=reduction,
                    dice_class_weight=class_weight_tensor(class_weight, pred.device), dice_loss_weight=1.0,
                    dice_ignore_index=ignore_index, dice_smooth=float(smooth), dice_exponent=float(exponent),
                    dice_avg_factor=None if avg_factor is None else float(avg_factor))
    _, loss, _ = run_fused(pred, target, None, spec)
    return _match_dtype(loss, pred)


class DiceLoss(nn.Module):
    """Drop-in for the reference's ``DiceLoss`` (:61-148)."""

    def __init__(self, smooth=1, exponent=2, reduction='mean', class_weight=None, loss_weight=1.0, ignore_index=255,
                 loss_name='loss_dice', **kwargs):
        super().__init__()
        self.smooth = smooth
        self.exponent = exponent
        self.reduction = reduction
        self.class_weight = get_class_weight(class_weight)
        self.loss_weight = loss_weight
        self.ignore_index = ignore_index
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

    def spec(self, device, avg_factor=None, reduction_override=None):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if avg_factor is not None and reduction == 'sum':
            raise ValueError('avg_factor can not be used with reduction="sum"')
        return LossSpec(want_dice=True, dice_reduction=reduction, dice_class_weight=self._class_weight_on(device),
                        dice_loss_weight=float(self.loss_weight), dice_ignore_index=self.ignore_index,
                        dice_smooth=float(self.smooth), dice_exponent=float(self.exponent),
                        dice_avg_factor=None if avg_factor is None else float(avg_factor))

    def forward(self, pred, target, avg_factor=None, reduction_override=None, **kwargs):
        spec = self.spec(pred.device, avg_factor, reduction_override)
        _, loss, _ = run_fused(pred, target, None, spec)
        return _match_dtype(loss, pred)

    @property
    def loss_name(self):
        return self._loss_name
