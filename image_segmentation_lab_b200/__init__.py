"""image_segmentation_lab_b200 — the per-pixel logits -> loss -> metrics hot path of
HanHan-TR/Image_Segmentation_lab as hand-written sm_100a CUDA behind the reference's Python signatures.

    resize, Upsample, add_prefix          utils/ops.py
    CrossEntropyLoss, cross_entropy       models/losses/cross_entropy_loss.py
    DiceLoss                              models/losses/dice_loss.py
    TverskyLoss                           models/losses/tversky_loss.py
    LovaszLoss                            models/losses/lovasz_loss.py
    accuracy, Accuracy                    models/losses/accuracy.py
    SegEvaluator                          core/evaluation/metrics.py
    fused_resize_losses, B200DecodeHeadLossMixin   models/decode_heads/decode_head.py:261-321 (fused)
    registry.install(LOSS)                registry/register.py, models/builder.py:40,262-283
    parse_losses                          utils/train_utils.py:31-74 (one all-reduce + one D2H per step)

Everything computes in libb200seg.so (include/b200seg.h); there is no CPU or PyTorch fallback.
"""
from . import distributed, registry
from ._lib import launch_count, lib_path, load as load_library
from .evaluation import ImageTable, SegEvaluator, area_totals_device, areas_device, prepare_images
from .fused import B200DecodeHeadLossMixin, fused_resize_losses
from .losses import (Accuracy, CrossEntropyLoss, DiceLoss, LovaszLoss, TverskyLoss, accuracy, binary_cross_entropy, cross_entropy, dice_loss, get_class_weight,
                     reduce_loss, weight_reduce_loss, weighted_loss)
from .ops import Upsample, add_prefix, resize
from .train_utils import parse_losses

__version__ = '0.1.0'

__all__ = [
    'resize', 'Upsample', 'add_prefix', 'CrossEntropyLoss', 'cross_entropy', 'binary_cross_entropy', 'DiceLoss', 'dice_loss', 'TverskyLoss', 'LovaszLoss', 'accuracy',
    'Accuracy', 'SegEvaluator', 'areas_device', 'area_totals_device', 'prepare_images', 'ImageTable', 'fused_resize_losses', 'B200DecodeHeadLossMixin', 'registry',
    'distributed', 'get_class_weight', 'reduce_loss', 'weight_reduce_loss', 'weighted_loss', 'load_library',
    'lib_path', 'launch_count', 'parse_losses',
]
