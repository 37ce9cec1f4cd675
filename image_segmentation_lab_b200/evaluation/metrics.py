"""SegEvaluator with the reference's interface (core/evaluation/metrics.py:25-356) on csrc/confusion.cu.

What changes underneath:
  * one kernel launch per ``process`` / ``intersect_and_union`` call for the whole list of images, fused
    arg-max + three area histograms, exact int64 counts accumulated on the device;
  * no host synchronisation while batches are processed — device results are queued and read back once,
    when ``results`` or ``compute_metrics`` is first used;
  * totals are summed in int64 (the reference sums fp32 tensors, which drops counts above 2**24; pass
    ``exact_totals=False`` to reproduce that rounding).
What is not provided: ``plot_results`` (cv2/PIL visualisation, out of scope) — ``show_result`` is accepted
and ignored with a warning.
"""
import warnings
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import _lib


def _device_of(tensors):
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    return torch.device('cuda', torch.cuda.current_device())


def _common(tensors, device, allowed, fallback):
    """Contiguous CUDA tensors of one dtype the kernel reads directly."""
    dts = {t.dtype for t in tensors}
    dt = dts.pop() if len(dts) == 1 else None
    if dt is None or dt not in allowed:
        dt = fallback
    if all(t.dtype == dt and t.device == device and t.is_contiguous() for t in tensors):
        return tensors, dt            # the usual case: nothing to convert, no per-tensor dispatch
    out = []
    for t in tensors:
        if t.device != device or t.dtype != dt:
            t = t.to(device=device, dtype=dt, non_blocking=True)
        out.append(t.contiguous())
    return out, dt


class ImageTable:
    """Device-side description of a list of (prediction, ground truth) images: the per-image table
    {pred*, gt*, n_pixels, (h,w), (H,W)} and the chunk prefix sum the kernels walk, built once and uploaded with one
    host-to-device copy. Build it with :func:`prepare_images` and pass it to :func:`areas_device` /
    :func:`area_totals_device` in place of the two lists when the same buffers are evaluated again (the table keeps
    the tensors alive)."""

    def __init__(self, preds, gts, from_logits, num_classes):
        lib = _lib.load()
        n = len(preds)
        self.n = n
        self.from_logits = bool(from_logits)
        self.resized = False
        self.device = _device_of(list(preds) + list(gts)) if n else torch.device('cuda', torch.cuda.current_device())
        self.chunk = lib.b200seg_confusion_chunk_pixels()
        if n == 0:
            self.total_chunks, self.meta, self.gdt, self.pdt, self.gts, self.preds = 0, None, None, None, [], []
            return
        dev = self.device
        gts_c, self.gdt = _common(list(gts), dev, _lib.LABEL_DTYPES, torch.float32)
        if from_logits:
            preds_c, self.pdt = _common(list(preds), dev, _lib.LOGIT_DTYPES, torch.float32)
        else:
            preds_c, self.pdt = _common(list(preds), dev, _lib.LABEL_DTYPES, torch.int64)
        self.preds, self.gts = preds_c, gts_c
        # a few vectorised passes (a 500-image sweep must not be bound by this loop)
        table = np.zeros((n, 5), dtype=np.int64)
        table[:, 0] = np.fromiter((p.data_ptr() for p in preds_c), dtype=np.int64, count=n)
        table[:, 1] = np.fromiter((g.data_ptr() for g in gts_c), dtype=np.int64, count=n)
        npx = np.fromiter((g.numel() for g in gts_c), dtype=np.int64, count=n)
        table[:, 2] = npx
        if from_logits:
            Cn = int(num_classes)
            for i, (p, g) in enumerate(zip(preds_c, gts_c)):
                if p.dim() == 4:
                    assert p.size(0) == 1, 'each prediction must be (1,C,H,W)'
                assert p.shape[-3] == Cn, 'logits have %d classes, evaluator has %d' % (p.shape[-3], Cn)
                if tuple(p.shape[-2:]) != tuple(g.shape[-2:]):
                    assert g.dim() >= 2, 'a 2-D ground truth is needed to resize the logits to it'
                    self.resized = True
                table[i, 3] = int(p.shape[-2]) | (int(p.shape[-1]) << 32)
                table[i, 4] = int(g.shape[-2]) | (int(g.shape[-1]) << 32)
        else:
            pnum = np.fromiter((p.numel() for p in preds_c), dtype=np.int64, count=n)
            assert (pnum == npx).all(), 'prediction / ground-truth size mismatch'
        prefix = np.zeros(n + 1, dtype=np.int64)
        np.cumsum((npx + self.chunk - 1) // self.chunk, out=prefix[1:])
        self.total_chunks = int(prefix[n])
        with torch.cuda.device(dev):
            self.meta = torch.from_numpy(np.concatenate([table.reshape(-1), prefix])).to(dev, non_blocking=True)
        self.images_p = self.meta.data_ptr()
        self.prefix_p = self.images_p + table.size * 8


def prepare_images(preds: Sequence[torch.Tensor], gts: Sequence[torch.Tensor], num_classes: int,
                   from_logits: bool = False) -> ImageTable:
    assert len(preds) == len(gts)  # metrics.py:236
    return ImageTable(preds, gts, from_logits, num_classes)


def areas_device(preds, gts: Optional[Sequence[torch.Tensor]], num_classes: int, ignore_index: int,
                 from_logits: bool = False, pred_maps: Optional[list] = None, align_corners: bool = False,
                 totals_only: bool = False) -> torch.Tensor:
    """int64 (n_images, 3, C) device tensor [intersect, pred, label] for a list of images; no host sync.
    With ``totals_only`` the result is the (3, C) sum over the images, accumulated inside the kernel.

    ``preds[i]`` is a label map (H_i,W_i) — or, with ``from_logits``, logits (1,C,H_i,W_i) / (C,H_i,W_i) whose
    arg-max over classes is taken in the same kernel (lowest index wins ties). ``gts[i]`` is (H_i,W_i) in
    any integer / float dtype (the reference's ``ori_gt`` is float32, core/dataset/kvasir_seg.py:37).
    ``preds`` may also be an :class:`ImageTable` from :func:`prepare_images` (then ``gts`` is ignored).
    ``pred_maps``: optional list that receives the int64 arg-max maps (logits mode only).
    Logits whose spatial size differs from their ground truth are bilinearly resized to it (``align_corners``)
    INSIDE the arg-max kernel — the rescale of decode_head.py:297-320 without materialising the (1,C,H,W) tensor.
    """
    lib = _lib.load()
    Cn = int(num_classes)
    if isinstance(preds, ImageTable):
        tab = preds
        from_logits = tab.from_logits
    else:
        assert len(preds) == len(gts)  # metrics.py:236
        tab = ImageTable(preds, gts, from_logits, Cn)
    n, dev = tab.n, tab.device
    shape = (3, Cn) if totals_only else (n, 3, Cn)
    if n == 0:
        return torch.zeros(shape, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        areas = torch.zeros(shape, dtype=torch.int64, device=dev)
        stream = _lib.stream_ptr(dev)
        tot = int(bool(totals_only))
        if from_logits:
            pout_p = None
            if pred_maps is not None:
                maps = [torch.empty(tuple(g.shape), dtype=torch.int64, device=dev) for g in tab.gts]
                ptrs = torch.tensor([m.data_ptr() for m in maps], dtype=torch.int64).to(dev, non_blocking=True)
                pout_p = ptrs.data_ptr()
                pred_maps.extend(maps)
            if tab.resized:
                _lib.check(lib.b200seg_confusion_logits_resized(
                    tab.images_p, tab.prefix_p, n, tab.total_chunks, tab.chunk, _lib.LOGIT_DTYPES[tab.pdt],
                    _lib.LABEL_DTYPES[tab.gdt], Cn, int(ignore_index), int(bool(align_corners)), areas.data_ptr(), pout_p, tot,
                    stream))
            else:
                _lib.check(lib.b200seg_confusion_logits(tab.images_p, tab.prefix_p, n, tab.total_chunks, tab.chunk,
                                                        _lib.LOGIT_DTYPES[tab.pdt], _lib.LABEL_DTYPES[tab.gdt], Cn,
                                                        int(ignore_index), areas.data_ptr(), pout_p, tot, stream))
        else:
            _lib.check(lib.b200seg_confusion_labels(tab.images_p, tab.prefix_p, n, tab.total_chunks, tab.chunk,
                                                    _lib.LABEL_DTYPES[tab.pdt], _lib.LABEL_DTYPES[tab.gdt], Cn,
                                                    int(ignore_index), areas.data_ptr(), tot, stream))
    return areas


def area_totals_device(preds, gts, num_classes: int, ignore_index: int, from_logits: bool = False,
                       align_corners: bool = False) -> torch.Tensor:
    """int64 (3, C) device totals [intersect, pred, label] over all images — what seg_metrics sums up (metrics.py:163-166)
    — accumulated inside the kernel: one launch, one (3, C) buffer, ready for a single int64 all-reduce across ranks."""
    return areas_device(preds, gts, num_classes, ignore_index, from_logits=from_logits, align_corners=align_corners,
                        totals_only=True)


def _iupl(areas: torch.Tensor) -> torch.Tensor:
    """(n,3,C) [I,P,L] -> (n,4,C) [I,U,P,L] (U = L + P - I, metrics.py:268)."""
    i, p, l = areas[:, 0], areas[:, 1], areas[:, 2]
    return torch.stack([i, l + p - i, p, l], dim=1)


class SegEvaluator():
    """IoU / Dice / F-score evaluator; constructor and methods as the reference's (metrics.py:52-356)."""

    def __init__(self,
                 epoch: int,
                 num_classes: int,
                 class_names: List[str],
                 palette: Sequence[Sequence[int]],
                 ignore_index: int = 255,
                 iou_metrics: List[str] = ['mIoU', 'mDice', 'mFscore'],
                 nan_to_num: Optional[int] = None,
                 beta: int = 1,
                 show_result: bool = True,
                 output_dir: Optional[str] = None,
                 format_only: bool = False,
                 prefix: Optional[str] = None,
                 exact_totals: bool = True,
                 keep_pred_maps: bool = False,
                 align_corners: bool = False,
                 **kwargs) -> None:
        self.epoch = epoch
        self.num_classes = num_classes
        self.class_names = class_names
        self.palette = palette
        self.ignore_index = ignore_index
        self.metrics = iou_metrics
        self.nan_to_num = nan_to_num
        self.beta = beta
        self.show_result = show_result
        self.output_dir = output_dir
        self.prefix = prefix
        if self.output_dir:
            import os
            os.makedirs(os.path.expanduser(self.output_dir), exist_ok=True)
        self.format_only = format_only
        self.exact_totals = exact_totals
        self.keep_pred_maps = keep_pred_maps
        self.align_corners = align_corners   # used when logits arrive at a lower resolution than the ground truth
        self._results = dict()   # key -> [[I...],[U...],[P...],[L...]] float32 CPU tensors (reference layout)
        self._pending = dict()   # key -> list of int64 (n,4,C) device tensors not yet read back
        self._exact = dict()     # key -> int64 (4,C) CPU running totals
        self._warned_plot = False

    # ------------------------------------------------------------------ reference-visible state
    @property
    def results(self):
        """key -> four lists of per-image (C,) float32 CPU tensors, as the reference stores them (:83,:121-124)."""
        self._drain()
        return self._results

    @results.setter
    def results(self, value):
        self._results = value
        self._pending = dict()
        self._exact = dict()

    def _drain(self):
        if not self._pending:
            return
        pending, self._pending = self._pending, dict()
        for key, chunks in pending.items():
            host = torch.cat(chunks, dim=0).cpu()  # one device->host copy per head
            lists = self._results.setdefault(key, [[], [], [], []])
            as_f32 = host.to(torch.float32)
            for j in range(4):
                lists[j].extend(as_f32[:, j].unbind(0))
            tot = host.sum(dim=0)
            self._exact[key] = self._exact[key] + tot if key in self._exact else tot

    def area_totals(self, key=None):
        """Exact int64 (4,C) totals [intersect, union, pred, label] for ``key`` (or a dict for all keys)."""
        self._drain()
        return self._exact if key is None else self._exact[key]

    def areas_on_device(self):
        """key -> int64 (4,C) DEVICE totals of what has been processed so far, without a host sync
        (for an NCCL all-reduce across ranks before the read-back)."""
        out = {}
        for key, chunks in self._pending.items():
            out[key] = torch.cat(chunks, dim=0).sum(dim=0)
        return out

    # ------------------------------------------------------------------ per batch
    def process(self, batch_idx: int, pred_batch, batch_infos: dict) -> None:
        """One validation batch (reference :85-124): ``pred_batch`` maps head name -> list of (1,C,H_i,W_i)
        logits; ``batch_infos['ori_gt']`` is the list of (H_i,W_i) ground-truth maps."""
        labels_batch = batch_infos['ori_gt']
        if self.num_classes == 1:
            raise NotImplementedError('num_classes == 1 (sigmoid) evaluation degenerates in the reference '
                                      '(argmax over one channel, histc with min == max); use num_classes=2')
        if self.show_result and batch_idx < 4 and not self._warned_plot:
            warnings.warn('SegEvaluator.plot_results (cv2 visualisation) is outside the B200 hot path; skipped')
            self._warned_plot = True
        for key, value in pred_batch.items():
            preds = [value[i] for i in range(len(value))]
            maps = [] if self.keep_pred_maps else None
            areas = areas_device(preds, list(labels_batch), self.num_classes, self.ignore_index, from_logits=True,
                                 pred_maps=maps, align_corners=self.align_corners)
            if maps is not None and isinstance(value, list):
                for i, m in enumerate(maps):
                    value[i] = m  # the reference replaces the logits by the label maps in place (:107)
            self._pending.setdefault(key, []).append(_iupl(areas))

    def compute_metrics(self):
        results = self.results
        if isinstance(results, list):
            return self.seg_metrics(results)
        if isinstance(results, dict):
            metrics_results = dict()
            for key, value in results.items():
                assert isinstance(value, list), "the values in the results dict of SegEvaluator must be a list"
                print('-------------------------' + key + '-------------------------')
                metrics_results[key] = self.seg_metrics(value, _exact=self._exact.get(key))
            return metrics_results
        raise TypeError('the results of SegEvaluator must be a list or dict')

    def seg_metrics(self, results: list, _exact=None) -> Dict[str, float]:
        """aAcc / mIoU / mAcc / mDice / mFscore / mPrecision / mRecall + per-class arrays (reference :139-208)."""
        assert len(results) == 4
        if self.exact_totals:
            if _exact is None:
                _exact = torch.stack([torch.stack([t.to(torch.int64) for t in results[j]]).sum(0) for j in range(4)])
            totals = [_exact[j].to(torch.float32) for j in range(4)]
        else:
            totals = [sum(results[j]) for j in range(4)]  # fp32 running sum, as the reference (:163-166)
        ret_metrics = self.total_area_to_metrics(totals[0], totals[1], totals[2], totals[3], self.metrics,
                                                 self.nan_to_num, self.beta)
        summary = OrderedDict((name, np.round(np.nanmean(val) * 100, 2)) for name, val in ret_metrics.items())
        metrics = dict()
        for key, val in summary.items():
            metrics[key if key == 'aAcc' else 'm' + key] = val
        ret_metrics.pop('aAcc', None)
        per_class = OrderedDict((name, np.round(val * 100, 2)) for name, val in ret_metrics.items())
        per_class.update({'Class': self.class_names})
        per_class.move_to_end('Class', last=False)
        print('\n' + self._format_table(per_class))
        metrics.update(per_class)
        return metrics

    def _format_table(self, columns):
        try:
            from prettytable import PrettyTable
            table = PrettyTable()
            for key, val in columns.items():
                table.add_column(key, val)
            return table.get_string()
        except ImportError:
            keys = list(columns.keys())
            rows = [' | '.join('%12s' % k for k in keys)]
            n = len(columns[keys[0]])
            for r in range(n):
                rows.append(' | '.join('%12s' % (columns[k][r],) for k in keys))
            return '\n'.join(rows)

    # ------------------------------------------------------------------ static API of the reference
    @staticmethod
    def intersect_and_union(pred_labels: list, labels_gt: list, num_classes: int, ignore_index: int):
        """Per-image areas (reference :210-270): four lists (intersect, union, pred, label) of (C,) float32
        CPU tensors. Unlike the reference the input lists are left untouched and there is a single
        device->host copy for the whole call. Use :func:`areas_device` to stay on the GPU."""
        assert len(pred_labels) == len(labels_gt)
        areas = _iupl(areas_device(pred_labels, labels_gt, num_classes, ignore_index)).cpu().to(torch.float32)
        return (list(areas[:, 0].unbind(0)), list(areas[:, 1].unbind(0)), list(areas[:, 2].unbind(0)),
                list(areas[:, 3].unbind(0)))

    @staticmethod
    def total_area_to_metrics(total_area_intersect, total_area_union, total_area_pred_label, total_area_label,
                              metrics: List[str] = ['mIoU'], nan_to_num: Optional[int] = None, beta: int = 1):
        """Per-class metrics from the four area totals (reference :272-356); fp32 tensor arithmetic on the
        host, 0/0 -> NaN unless ``nan_to_num``. Raises KeyError for an unsupported metric (:319-320)."""
        if isinstance(metrics, str):
            metrics = [metrics]
        allowed_metrics = ['mIoU', 'mDice', 'mFscore']
        if not set(metrics).issubset(set(allowed_metrics)):
            raise KeyError(f'metrics {metrics} is not supported')
        I, U = torch.as_tensor(total_area_intersect), torch.as_tensor(total_area_union)
        P, L = torch.as_tensor(total_area_pred_label), torch.as_tensor(total_area_label)
        ret = OrderedDict({'aAcc': I.sum() / L.sum()})
        for metric in metrics:
            if metric == 'mIoU':
                ret['IoU'] = I / U
                ret['Acc'] = I / L
            elif metric == 'mDice':
                ret['Dice'] = 2 * I / (P + L)
                ret['Acc'] = I / L
            elif metric == 'mFscore':
                precision = I / P
                recall = I / L
                ret['Fscore'] = (1 + beta ** 2) * (precision * recall) / ((beta ** 2 * precision) + recall)
                ret['Precision'] = precision
                ret['Recall'] = recall
        ret = {name: value.numpy() for name, value in ret.items()}
        if nan_to_num is not None:
            ret = OrderedDict({name: np.nan_to_num(value, nan=nan_to_num) for name, value in ret.items()})
        return ret

    def plot_results(self, *args, **kwargs):
        raise NotImplementedError('plot_results (cv2/PIL visualisation) is outside the B200 hot path')
