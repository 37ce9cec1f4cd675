"""ctypes binding of libb200seg.so (see include/b200seg.h).

The CUDA library is the only compute path of this package: if it cannot be loaded the package
raises — there is no CPU or PyTorch fallback.
"""
import ctypes as C
import os

import torch

from . import _build

_LIB = None

# enums (mirror include/b200seg.h)
F32, BF16, F16 = 0, 1, 2
L_U8, L_I16, L_I32, L_I64, L_F32, L_F64 = 0, 1, 2, 3, 4, 5
WANT_CE, WANT_DICE, WANT_ACC, WANT_LOSS_PX, WANT_LSE = 1, 2, 4, 8, 16
MODE_DICE, MODE_TVERSKY = 0, 1
ST_CE_SUM, ST_N_VALID, ST_N_CORRECT, ST_N_BAD, ST_N_ACC, STATS_WORDS = 0, 1, 2, 3, 4, 8
RED_NONE, RED_MEAN, RED_SUM = 0, 1, 2
ABI_VERSION = 14
LOG_CE_SUM, LOG_N_VALID, LOG_N_CORRECT, LOG_N_ACC, LOG_N_BAD, LOG_N_PIXELS, LOG_DICE_SUM, LOG_N_IMAGES, LOG_WORDS = \
    0, 1, 2, 3, 4, 5, 6, 7, 8

LOGIT_DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}
LABEL_DTYPES = {
    torch.uint8: L_U8, torch.int16: L_I16, torch.int32: L_I32, torch.int64: L_I64,
    torch.float32: L_F32, torch.float64: L_F64,
}
REDUCTIONS = {"none": RED_NONE, "mean": RED_MEAN, "sum": RED_SUM}


class LossDesc(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("labels", C.c_void_p), ("pixel_weight", C.c_void_p), ("ce_class_weight", C.c_void_p),
        ("logit_dtype", C.c_int32), ("label_dtype", C.c_int32),
        ("N", C.c_int32), ("C", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("align_corners", C.c_int32), ("flags", C.c_int32),
        ("ignore_index", C.c_int64),
        ("acc_has_ignore", C.c_int32), ("dice_mode", C.c_int32),
        ("acc_ignore_index", C.c_int64),
        ("dice_ignore_index", C.c_int64),
        ("dice_exponent", C.c_float), ("reserved1", C.c_float),
        ("lse", C.c_void_p), ("loss_px", C.c_void_p),
        ("ce_loss_weight", C.c_float), ("reserved2", C.c_float),
        ("stats", C.c_void_p), ("dice_part", C.c_void_p),
    ]


class FinalizeDesc(C.Structure):
    _fields_ = [
        ("stats", C.c_void_p), ("dice_part", C.c_void_p), ("dice_class_weight", C.c_void_p),
        ("N", C.c_int32), ("C", C.c_int32),
        ("n_pixels", C.c_int64),
        ("ce_reduction", C.c_int32), ("ce_avg_non_ignore", C.c_int32),
        ("ce_has_avg_factor", C.c_int32), ("dice_has_avg_factor", C.c_int32),
        ("ce_avg_factor", C.c_double), ("dice_avg_factor", C.c_double),
        ("ce_loss_weight", C.c_float), ("dice_loss_weight", C.c_float),
        ("dice_smooth", C.c_float), ("dice_reduction", C.c_int32),
        ("dice_ignore_index", C.c_int64),
        ("out_loss_ce", C.c_void_p), ("out_loss_dice", C.c_void_p), ("out_acc", C.c_void_p),
        ("dice_coef", C.c_void_p), ("log_vec", C.c_void_p),
        ("dice_mode", C.c_int32), ("tversky_alpha", C.c_float), ("tversky_beta", C.c_float), ("reserved0", C.c_int32),
    ]


class LossBwdDesc(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("labels", C.c_void_p), ("pixel_weight", C.c_void_p), ("ce_class_weight", C.c_void_p),
        ("lse", C.c_void_p),
        ("logit_dtype", C.c_int32), ("label_dtype", C.c_int32),
        ("N", C.c_int32), ("C", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("align_corners", C.c_int32), ("flags", C.c_int32),
        ("ignore_index", C.c_int64), ("dice_ignore_index", C.c_int64),
        ("dice_exponent", C.c_float), ("ce_scale_host", C.c_float),
        ("ce_grad_out", C.c_void_p), ("ce_grad_px", C.c_void_p), ("stats", C.c_void_p),
        ("ce_use_nvalid", C.c_int32), ("dice_mode", C.c_int32),
        ("dice_coef", C.c_void_p), ("dice_grad_out", C.c_void_p),
        ("grad_logits", C.c_void_p), ("reserved_scratch", C.c_void_p), ("scratch_px", C.c_void_p),
    ]


class LossFusedDesc(C.Structure):
    _fields_ = [
        ("fwd", LossDesc),
        ("grad_scale_host", C.c_float), ("use_nvalid", C.c_int32),
        ("grad_out", C.c_void_p), ("grad_logits", C.c_void_p), ("workspace", C.c_void_p),
        ("defer_combine", C.c_int32), ("reserved", C.c_int32),
    ]


class BceDesc(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("labels", C.c_void_p), ("pixel_weight", C.c_void_p), ("pos_weight", C.c_void_p),
        ("logit_dtype", C.c_int32), ("label_dtype", C.c_int32), ("N", C.c_int32), ("C", C.c_int32),
        ("HW", C.c_int64), ("ignore_index", C.c_int64),
        ("single_channel", C.c_int32), ("use_nvalid", C.c_int32),
        ("loss_weight", C.c_float), ("grad_scale_host", C.c_float),
        ("loss_elem", C.c_void_p), ("grad_out", C.c_void_p), ("grad_elem", C.c_void_p), ("grad_logits", C.c_void_p),
        ("stats", C.c_void_p), ("out", C.c_void_p), ("out_scale_host", C.c_float), ("acc_has_ignore", C.c_int32),
        ("acc_out", C.c_void_p), ("acc_ignore_index", C.c_int64),
    ]


class LovaszDesc(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("labels", C.c_void_p), ("lse", C.c_void_p), ("class_weight", C.c_void_p),
        ("logit_dtype", C.c_int32), ("label_dtype", C.c_int32), ("N", C.c_int32), ("C", C.c_int32),
        ("HW", C.c_int64), ("ignore_index", C.c_int64),
        ("has_ignore", C.c_int32), ("binary", C.c_int32), ("per_image", C.c_int32), ("only_present", C.c_int32),
        ("classes_host", C.POINTER(C.c_int32)), ("n_classes", C.c_int32), ("reduction", C.c_int32),
        ("has_avg_factor", C.c_int32), ("loss_weight", C.c_float),
        ("avg_factor", C.c_double),
        ("lab16", C.c_void_p), ("G", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("seg_stats", C.c_void_p), ("out", C.c_void_p), ("coef", C.c_void_p),
    ]


class LovaszBwdDesc(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("lse", C.c_void_p), ("lab16", C.c_void_p), ("G", C.c_void_p), ("coef", C.c_void_p),
        ("grad_out", C.c_void_p), ("grad_logits", C.c_void_p),
        ("logit_dtype", C.c_int32), ("N", C.c_int32), ("C", C.c_int32),
        ("binary", C.c_int32), ("per_image", C.c_int32), ("grad_per_group", C.c_int32),
        ("HW", C.c_int64),
    ]


class Image(C.Structure):
    _fields_ = [
        ("pred", C.c_void_p), ("gt", C.c_void_p), ("n_pixels", C.c_int64),
        ("h", C.c_int32), ("w", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
    ]


# every symbol include/b200seg.h declares: (name, restype, argtypes)
_i32, _i64, _f, _p = C.c_int32, C.c_int64, C.c_float, C.c_void_p
SYMBOLS = [
    ("b200seg_loss_fwd", C.c_int, [C.POINTER(LossDesc), _p]),
    ("b200seg_loss_finalize", C.c_int, [C.POINTER(FinalizeDesc), _p]),
    ("b200seg_loss_bwd", C.c_int, [C.POINTER(LossBwdDesc), _p]),
    ("b200seg_loss_fused_workspace_bytes", _i64, [_i32] * 7),
    ("b200seg_loss_flat_single_ok", _i32, [_p, _p, _i32, _i32, _i32, _i64, _i32]),
    ("b200seg_loss_fused_fwdbwd", C.c_int, [C.POINTER(LossFusedDesc), _p]),
    ("b200seg_loss_fused_combine", C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _f, _p, _i32, _p, _p]),
    ("b200seg_scale_inplace", C.c_int, [_p, _i32, _i64, _p, _p]),
    ("b200seg_resize_bilinear_fwd", C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f, _f, _p]),
    ("b200seg_resize_bilinear_bwd", C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f, _f, _p]),
    ("b200seg_resize_nearest_fwd", C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _f, _f, _p]),
    ("b200seg_resize_nearest_bwd", C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _f, _f, _p]),
    ("b200seg_confusion_labels", C.c_int, [_p, _p, _i32, _i64, _i32, _i32, _i32, _i32, _i64, _p, _i32, _p]),
    ("b200seg_confusion_logits", C.c_int, [_p, _p, _i32, _i64, _i32, _i32, _i32, _i32, _i64, _p, _p, _i32, _p]),
    ("b200seg_confusion_logits_resized", C.c_int, [_p, _p, _i32, _i64, _i32, _i32, _i32, _i32, _i64, _i32, _p, _p, _i32, _p]),
    ("b200seg_confusion_chunk_pixels", _i32, []),
    ("b200seg_topk_counts", C.c_int,
     [_p, _p, _i32, _i32, _i32, _i32, _i64, _i32, _i64, C.POINTER(_i32), _i32, _i32, _f, _p, _p]),
    ("b200seg_bce_fwd", C.c_int, [C.POINTER(BceDesc), _p]),
    ("b200seg_bce_bwd", C.c_int, [C.POINTER(BceDesc), _p]),
    ("b200seg_lovasz_workspace_bytes", _i64, [_i32, _i32, _i64, _i32, _i32]),
    ("b200seg_lovasz_fwd", C.c_int, [C.POINTER(LovaszDesc), _p]),
    ("b200seg_lovasz_bwd", C.c_int, [C.POINTER(LovaszBwdDesc), _p]),
    ("b200seg_last_error", C.c_char_p, []),
    ("b200seg_abi_version", _i32, []),
    ("b200seg_launch_count", _i64, []),
]


def lib_path():
    return _build.LIB_PATH


def load():
    """Load libb200seg.so (building it first if nvcc is available and it is missing/stale)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    alt = os.environ.get("B200SEG_LIB_PATH")     # an alternative build of the same sources (A/B measurements of compile-time variants)
    if alt:
        path = alt
    elif not os.path.exists(path) or (not _build.is_current() and os.environ.get("B200SEG_NO_REBUILD") != "1"):
        try:
            _build.build_library()
        except Exception as e:  # no nvcc: use a prebuilt .so if there is one, else fail loudly
            if not os.path.exists(path):
                raise RuntimeError(
                    "libb200seg.so is missing and could not be built (%s). This package has no CPU fallback: "
                    "run `python -c 'import __graft_entry__ as g; g.build()'` on a machine with nvcc." % (e,)) from e
    lib = C.CDLL(path)
    for name, restype, argtypes in SYMBOLS:
        if not hasattr(lib, name):
            raise RuntimeError("libb200seg.so does not export %s — rebuild it" % name)
        fn = getattr(lib, name)
        fn.restype = restype
        if argtypes is not None:
            fn.argtypes = argtypes
    v = lib.b200seg_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError("libb200seg.so ABI version %d != expected %d — rebuild it" % (v, ABI_VERSION))
    _LIB = lib
    return lib


def last_error():
    return load().b200seg_last_error().decode("utf-8", "replace")


def check(rc, exc=RuntimeError):
    if rc != 0:
        msg = last_error()
        if "avg_factor can not be used" in msg:
            raise ValueError(msg)  # models/losses/utils.py:78-79
        raise exc(msg)


def launch_count():
    return int(load().b200seg_launch_count())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            "%s must be a CUDA tensor: image_segmentation_lab_b200 runs on B200 (sm_100a) only and has no CPU path "
            "(got device %s)" % (what, t.device))


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
