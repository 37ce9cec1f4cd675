"""TverskyLoss with the reference's signatures (models/losses/tversky_loss.py:24-148) on the streaming overlap-loss
kernels (csrc/loss_dice.cu in Tversky mode: sums masked by ``target != ignore_index``, exponent 1).

Per sample n and class c != ignore_index (:52-68, on soft-max probabilities p, one-hot t of the labels clamped to
[0, C-1] and the valid mask v):  TP = sum p t v,  FP = sum p (1 - t) v,  FN = sum (1 - p) t v,
tversky = (TP + smooth) / (TP + alpha FP + beta FN + smooth); the loss is
``loss_weight * sum_c class_weight[c] * mean_n(1 - tversky[n, c]) / C`` (:24-49, both helpers are @weighted_loss with
the default 'mean'). ``forward`` swallows call-site kwargs (``weight=``, ``ignore_index=``) as the reference does (:112).
"""
import torch.nn as nn

from ._function import LossSpec, run_fused
from .cross_entropy_loss import _match_dtype
from .utils import class_weight_tensor, get_class_weight


class TverskyLoss(nn.Module):
    """Drop-in for the reference's ``TverskyLoss`` (:71-148)."""

    def __init__(self, smooth=1, class_weight=None, loss_weight=1.0, ignore_index=255, alpha=0.3, beta=0.7,
                 loss_name='loss_tversky'):
        super().__init__()
        self.smooth = smooth
        self.class_weight = get_class_weight(class_weight)
        self.loss_weight = loss_weight
        self.ignore_index = ignore_index
        assert (alpha + beta == 1.0), 'Sum of alpha and beta but be 1.0!'   # tversky_loss.py:106
        self.alpha = alpha
        self.beta = beta
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

    def spec(self, device):
        return LossSpec(want_dice=True, dice_mode='tversky', dice_reduction='mean',
                        dice_class_weight=self._class_weight_on(device), dice_loss_weight=float(self.loss_weight),
                        dice_ignore_index=self.ignore_index, dice_smooth=float(self.smooth), dice_exponent=1.0,
                        tversky_alpha=float(self.alpha), tversky_beta=float(self.beta))

    def forward(self, pred, target, **kwargs):
        _, loss, _ = run_fused(pred, target, None, self.spec(pred.device))
        return _match_dtype(loss, pred)

    @property
    def loss_name(self):
        return self._loss_name
