"""Top-k pixel accuracy with the reference's signature (models/losses/accuracy.py:6-92)."""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib
from ._function import LossSpec, prep_labels, run_fused

_EPS = float(torch.finfo(torch.float32).eps)


def accuracy(pred, target, topk=1, thresh=None, ignore_index=None):
    """Accuracy in percent, shape (1,) per k (tuple ``topk`` -> list), reference :6-61.

    ``100 * (correct + eps) / (n_valid + eps)`` over pixels with ``target != ignore_index`` (:51-60).
    top-1 without a threshold comes from the same pass as the loss (arg-max of the soft-max reduction);
    other k / thresh use the rank of the label's logit (b200seg_topk_counts). Ties rank the lower class
    index first (torch.topk leaves tie order unspecified).
    """
    assert isinstance(topk, (int, tuple))
    if isinstance(topk, int):
        topk = (topk,)
        return_single = True
    else:
        return_single = False
    maxk = max(topk)
    if pred.size(0) == 0:
        accu = [pred.new_tensor(0.) for _ in range(len(topk))]
        return accu[0] if return_single else accu
    assert pred.ndim == target.ndim + 1
    assert pred.size(0) == target.size(0)
    assert maxk <= pred.size(1), f'maxk {maxk} exceeds pred dimension {pred.size(1)}'
    _lib.require_cuda(pred, 'pred')
    from .cross_entropy_loss import _as_image
    pred4, lab, _, _ = _as_image(pred.detach(), target, None)
    if topk == (1,) and thresh is None:
        spec = LossSpec(want_acc=True, acc_ignore_index=ignore_index)
        _, _, acc = run_fused(pred4, lab, None, spec)
        res = [acc]  # float32 (1,), as correct.float().sum(0, keepdim=True) in the reference
        return res[0] if return_single else res
    if len(topk) > 4:
        raise NotImplementedError('at most 4 values of k per call')
    lib = _lib.load()
    pred4 = pred4.contiguous()
    if pred4.dtype not in _lib.LOGIT_DTYPES:
        raise TypeError('pred must be float32, bfloat16 or float16')
    lab = prep_labels(lab, pred4)
    N, Cc = pred4.shape[:2]
    HW = pred4.shape[2] * pred4.shape[3]
    counts = torch.empty(len(topk) + 1, dtype=torch.int64, device=pred.device)
    ks = (C.c_int32 * len(topk))(*[int(k) for k in topk])
    with torch.cuda.device(pred.device):
        _lib.check(lib.b200seg_topk_counts(
            pred4.data_ptr(), lab.data_ptr(), _lib.LOGIT_DTYPES[pred4.dtype], _lib.LABEL_DTYPES[lab.dtype], N, Cc, HW,
            int(ignore_index is not None), int(ignore_index or 0), ks, len(topk), int(thresh is not None),
            C.c_float(float(thresh or 0.0)), counts.data_ptr(), _lib.stream_ptr(pred.device)))
    cf = counts.to(torch.float64)
    total = cf[len(topk)] + _EPS
    res = [(cf[j:j + 1].to(torch.float32) + _EPS) * (100.0 / total).to(torch.float32) for j in range(len(topk))]
    return res[0] if return_single else res


class Accuracy(nn.Module):
    """Module form (reference :64-92)."""

    def __init__(self, topk=(1,), thresh=None, ignore_index=None):
        super().__init__()
        self.topk = topk
        self.thresh = thresh
        self.ignore_index = ignore_index

    def forward(self, pred, target):
        return accuracy(pred, target, self.topk, self.thresh, self.ignore_index)
