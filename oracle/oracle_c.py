"""ctypes wrapper of oracle/oracle_c.c (test infrastructure only — see the header of that file)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def load():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, 'liboracle_c.so')
        src = os.path.join(HERE, 'oracle_c.c')
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(['make', '-s', '-C', HERE], check=True)
        _LIB = C.CDLL(so)
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def ce(logits, labels, size, pixel_weight=None, class_weight=None, align_corners=False, ignore_index=-100,
       reduction='mean', avg_non_ignore=False, avg_factor=None, loss_weight=1.0, acc_ignore_index=None, grad_px=None):
    """Returns dict(loss, grad, n_valid, n_correct, n_acc, acc) in float64."""
    lib = load()
    x = _f32(logits)
    y = np.ascontiguousarray(labels, dtype=np.int64)
    N, Cc, h, w = x.shape
    H, W = size
    red = {'none': 0, 'mean': 1, 'sum': 2}[reduction]
    pw, cw = _f32(pixel_weight), _f32(class_weight)
    gpx = None if grad_px is None else np.ascontiguousarray(grad_px, dtype=np.float64)
    loss = np.zeros(1, dtype=np.float64)
    loss_px = np.zeros((N, H, W), dtype=np.float64) if red == 0 else None
    grad = np.zeros((N, Cc, h, w), dtype=np.float64)
    counts = np.zeros(3, dtype=np.int64)
    lib.oc_ce(_p(x, C.c_float), _p(y, C.c_int64), _p(pw, C.c_float), _p(cw, C.c_float), N, Cc, h, w, H, W,
              int(bool(align_corners)), C.c_int64(int(ignore_index)), red, int(bool(avg_non_ignore)),
              C.c_double(-1.0 if avg_factor is None else float(avg_factor)), C.c_double(float(loss_weight)),
              int(acc_ignore_index is not None), C.c_int64(int(acc_ignore_index or 0)), _p(gpx, C.c_double),
              _p(loss, C.c_double), _p(loss_px, C.c_double), _p(grad, C.c_double), _p(counts, C.c_int64))
    eps = float(np.finfo(np.float32).eps)
    return dict(loss=loss_px if red == 0 else loss[0], grad=grad, n_valid=int(counts[0]), n_correct=int(counts[1]),
                n_acc=int(counts[2]), acc=100.0 * (counts[1] + eps) / (counts[2] + eps))


def dice(logits, labels, class_weight=None, ignore_index=255, smooth=1.0, exponent=2.0, loss_weight=1.0, reduction='mean',
         avg_factor=None):
    lib = load()
    x = _f32(logits)
    y = np.ascontiguousarray(labels, dtype=np.int64)
    N, Cc, H, W = x.shape
    assert Cc <= 4096
    cw = _f32(class_weight)
    loss = np.zeros(1, dtype=np.float64)
    grad = np.zeros((N, Cc, H, W), dtype=np.float64)
    lib.oc_dice(_p(x, C.c_float), _p(y, C.c_int64), _p(cw, C.c_float), N, Cc, H, W, C.c_int64(int(ignore_index)),
                C.c_double(float(smooth)), C.c_double(float(exponent)), C.c_double(float(loss_weight)),
                {'none': 0, 'mean': 1, 'sum': 2}[reduction], C.c_double(-1.0 if avg_factor is None else float(avg_factor)),
                _p(loss, C.c_double), _p(grad, C.c_double))
    return dict(loss=loss[0], grad=grad)


def resize_bilinear(x, size, align_corners=False):
    lib = load()
    x = _f32(x)
    N, Cc, h, w = x.shape
    out = np.zeros((N, Cc, size[0], size[1]), dtype=np.float64)
    lib.oc_resize_bilinear(_p(x, C.c_float), _p(out, C.c_double), N * Cc, h, w, size[0], size[1], int(bool(align_corners)))
    return out


def argmax(logits):
    """logits (C,H,W) or (1,C,H,W) -> int64 (H,W)"""
    lib = load()
    x = _f32(logits)
    if x.ndim == 4:
        x = x[0]
    Cc, H, W = x.shape
    out = np.zeros((H, W), dtype=np.int64)
    lib.oc_argmax(_p(x, C.c_float), Cc, C.c_int64(H * W), _p(out, C.c_int64))
    return out


def areas(pred, gt, num_classes, ignore_index):
    """int64 (3,C): intersect, pred, label for one image."""
    lib = load()
    p = np.ascontiguousarray(pred, dtype=np.int64).reshape(-1)
    g = np.ascontiguousarray(gt, dtype=np.float32).reshape(-1)
    out = np.zeros(3 * num_classes, dtype=np.int64)
    lib.oc_areas(_p(p, C.c_int64), _p(g, C.c_float), C.c_int64(p.size), num_classes, C.c_int64(int(ignore_index)),
                 _p(out, C.c_int64))
    return out.reshape(3, num_classes)


def lovasz(logits, labels, loss_type='multi_class', classes='present', per_image=False, reduction='mean', class_weight=None,
           loss_weight=1.0, avg_factor=None, ignore_index=255, grad_out=None):
    """LovaszLoss forward + gradient in float64 (exact Jaccard increments). Returns dict(loss, grad)."""
    lib = load()
    x = _f32(logits)
    y = np.ascontiguousarray(labels, dtype=np.int64)
    binary = loss_type == 'binary'
    N = x.shape[0]
    Cc = 1 if binary else x.shape[1]
    HW = x.size // (N * Cc)
    mask = np.ones(Cc, dtype=np.uint8)
    if isinstance(classes, (list, tuple)):
        mask[:] = 0
        mask[list(classes)] = 1
    cw = _f32(class_weight)
    n_out = N if (per_image and reduction == 'none') else 1
    loss = np.zeros(n_out, dtype=np.float64)
    grad = np.zeros(x.shape, dtype=np.float64)
    go = None if grad_out is None else np.ascontiguousarray(grad_out, dtype=np.float64).reshape(-1)
    lib.oc_lovasz(_p(x, C.c_float), _p(y, C.c_int64), _p(cw, C.c_float), _p(mask, C.c_uint8), N, Cc, C.c_int64(HW),
                  int(ignore_index is not None), C.c_int64(int(ignore_index or 0)), int(binary), int(bool(per_image)),
                  int(classes == 'present'), {'none': 0, 'mean': 1, 'sum': 2}[reduction],
                  C.c_double(-1.0 if avg_factor is None else float(avg_factor)), C.c_double(float(loss_weight)),
                  _p(go, C.c_double), _p(loss, C.c_double), _p(grad, C.c_double))
    return dict(loss=loss if n_out > 1 else loss[0], grad=grad)
