from .metrics import SegEvaluator, areas_device

__all__ = ['SegEvaluator', 'areas_device']
