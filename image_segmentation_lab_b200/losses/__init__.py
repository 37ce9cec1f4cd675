from ._bce import binary_cross_entropy
from .accuracy import Accuracy, accuracy
from .cross_entropy_loss import CrossEntropyLoss, cross_entropy
from .dice_loss import DiceLoss, dice_loss
from .lovasz_loss import LovaszLoss
from .tversky_loss import TverskyLoss
from .utils import get_class_weight, reduce_loss, weight_reduce_loss, weighted_loss

__all__ = ['binary_cross_entropy', 'accuracy', 'Accuracy', 'cross_entropy', 'CrossEntropyLoss', 'dice_loss', 'DiceLoss', 'TverskyLoss', 'LovaszLoss', 'get_class_weight',
           'reduce_loss', 'weight_reduce_loss', 'weighted_loss']
