"""Reduction helpers with the reference's signatures (models/losses/utils.py:10-126).

These are host-side conveniences kept for API compatibility (custom losses built with
``@weighted_loss``); the fused kernels implement the same reduction rules on the device
(csrc/loss_stream.cu finalize_kernel) and never call them on the hot path.
"""
import functools

import numpy as np
import torch

_EPS = torch.finfo(torch.float32).eps


def get_class_weight(class_weight):
    """list passthrough; '.npy' via numpy; json/yaml/pkl by extension (reference :10-25 uses mmcv.load)."""
    if isinstance(class_weight, str):
        if class_weight.endswith(".npy"):
            class_weight = np.load(class_weight)
        elif class_weight.endswith(".json"):
            import json
            with open(class_weight) as fh:
                class_weight = json.load(fh)
        elif class_weight.endswith((".yaml", ".yml")):
            import yaml
            with open(class_weight) as fh:
                class_weight = yaml.safe_load(fh)
        elif class_weight.endswith((".pkl", ".pickle")):
            import pickle
            with open(class_weight, "rb") as fh:
                class_weight = pickle.load(fh)
        else:
            raise TypeError("unsupported class_weight file: %s" % class_weight)
    return class_weight


def reduce_loss(loss, reduction):
    """'none' | 'mean' | 'sum' (reference :28-45)."""
    if reduction == "none":
        return loss
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    raise ValueError("%s is not a valid value for reduction" % reduction)


def weight_reduce_loss(loss, weight=None, reduction="mean", avg_factor=None):
    """Element-wise weight, then reduce (reference :48-80)."""
    if weight is not None:
        assert weight.dim() == loss.dim()
        if weight.dim() > 1:
            assert weight.size(1) == 1 or weight.size(1) == loss.size(1)
        loss = loss * weight
    if avg_factor is None:
        loss = reduce_loss(loss, reduction)
    elif reduction == "mean":
        loss = loss.sum() / (avg_factor + _EPS)
    elif reduction != "none":
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


def weighted_loss(loss_func):
    """Decorator adding (weight, reduction, avg_factor) to an element-wise loss (reference :83-126)."""

    @functools.wraps(loss_func)
    def wrapper(pred, target, weight=None, reduction="mean", avg_factor=None, **kwargs):
        loss = loss_func(pred, target, **kwargs)
        return weight_reduce_loss(loss, weight, reduction, avg_factor)

    return wrapper


def class_weight_tensor(class_weight, device):
    """fp32 (C,) device tensor from a list / ndarray / tensor (reference: cls_score.new_tensor(...) per call)."""
    if class_weight is None:
        return None
    if isinstance(class_weight, torch.Tensor):
        return class_weight.detach().to(device=device, dtype=torch.float32).contiguous()
    return torch.as_tensor(np.asarray(class_weight, dtype=np.float32), device=device).contiguous()
