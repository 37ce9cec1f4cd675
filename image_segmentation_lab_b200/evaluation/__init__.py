from .metrics import ImageTable, SegEvaluator, area_totals_device, areas_device, prepare_images

__all__ = ['SegEvaluator', 'areas_device', 'area_totals_device', 'prepare_images', 'ImageTable']
