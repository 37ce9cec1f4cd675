"""``resize`` / ``Upsample`` / ``add_prefix`` with the reference's signatures (utils/ops.py:7-69).

``resize`` is a thin wrapper over F.interpolate in the reference; here bilinear and nearest run on
csrc/resize.cu (forward bit-identical to ATen's CUDA kernels, deterministic gather backwards, size= or
scale_factor= including fractional factors). The decode-head hot call
(decode_head.py:266-269) should not use this at all — ``fused_resize_losses`` never materialises the
up-sampled logits — but validation rescaling and inference still need the real tensor.
"""
import warnings

import torch
import torch.nn as nn

from . import _lib


def _out_size(input, size, scale_factor):
    if size is not None and scale_factor is not None:
        raise ValueError('only one of size or scale_factor should be defined')
    if size is None and scale_factor is None:
        raise ValueError('either size or scale_factor should be defined')
    if size is not None:
        if isinstance(size, int):
            size = (size, size)
        return int(size[0]), int(size[1])
    if not isinstance(scale_factor, (tuple, list)):
        scale_factor = (scale_factor, scale_factor)
    import math
    return (int(math.floor(float(input.shape[2]) * float(scale_factor[0]))),
            int(math.floor(float(input.shape[3]) * float(scale_factor[1]))))


def _kernel_scales(scale_factor):
    """ATen's compute_scales_value: with scale_factor= (and recompute_scale_factor unset) the source-index scale is
    (float)(1.0 / scale_factor) instead of in / out. 0 = derive from the sizes."""
    if scale_factor is None:
        return 0.0, 0.0
    sf = scale_factor if isinstance(scale_factor, (tuple, list)) else (scale_factor, scale_factor)
    return float(torch.tensor(1.0 / float(sf[0]), dtype=torch.float32)), float(torch.tensor(1.0 / float(sf[1]), dtype=torch.float32))


class _Resize(torch.autograd.Function):
    """F.interpolate(mode='bilinear' | 'nearest') on csrc/resize.cu: the bilinear forward is bit-identical to ATen's CUDA
    kernel (FMA contraction pinned), both backwards are deterministic gathers (ATen scatters with atomicAdd)."""

    @staticmethod
    def forward(ctx, x, H, W, mode, align_corners, scales):
        lib = _lib.load()
        x = x.contiguous()
        N, Cc, h, w = x.shape
        out = torch.empty((N, Cc, H, W), dtype=x.dtype, device=x.device)
        sh, sw = scales
        with torch.cuda.device(x.device):
            if mode == 'bilinear':
                _lib.check(lib.b200seg_resize_bilinear_fwd(x.data_ptr(), out.data_ptr(), _lib.LOGIT_DTYPES[x.dtype], N * Cc, h, w, H,
                                                           W, int(bool(align_corners)), sh, sw, _lib.stream_ptr(x.device)))
            else:
                _lib.check(lib.b200seg_resize_nearest_fwd(x.data_ptr(), out.data_ptr(), _lib.LOGIT_DTYPES[x.dtype], N * Cc, h, w, H, W,
                                                          sh, sw, _lib.stream_ptr(x.device)))
        ctx.geom = (N, Cc, h, w, H, W, mode, bool(align_corners), sh, sw)
        return out

    @staticmethod
    def backward(ctx, go):
        lib = _lib.load()
        N, Cc, h, w, H, W, mode, ac, sh, sw = ctx.geom
        go = go.contiguous()
        gi = torch.empty((N, Cc, h, w), dtype=go.dtype, device=go.device)
        with torch.cuda.device(go.device):
            if mode == 'bilinear':
                _lib.check(lib.b200seg_resize_bilinear_bwd(go.data_ptr(), gi.data_ptr(), _lib.LOGIT_DTYPES[go.dtype], N * Cc, h, w, H,
                                                           W, int(ac), sh, sw, _lib.stream_ptr(go.device)))
            else:
                _lib.check(lib.b200seg_resize_nearest_bwd(go.data_ptr(), gi.data_ptr(), _lib.LOGIT_DTYPES[go.dtype], N * Cc, h, w, H, W,
                                                          sh, sw, _lib.stream_ptr(go.device)))
        return gi, None, None, None, None, None


def resize(input, size=None, scale_factor=None, mode='nearest', align_corners=None, warning=True):
    """Same arguments, warning and result as the reference (utils/ops.py:7-26) for mode 'bilinear' and 'nearest' (the
    modes the segmentation path uses: decode_head.py:266-269,301-318, encoder_decoder.py:94-97,247-251); other modes of
    F.interpolate (bicubic, area, ...) are not on that path and raise NotImplementedError."""
    if warning:
        if size is not None and align_corners:
            input_h, input_w = tuple(int(x) for x in input.shape[2:])
            output_h, output_w = tuple(int(x) for x in size)
            if output_h > input_h or output_w > output_h:
                if ((output_h > 1 and output_w > 1 and input_h > 1 and input_w > 1) and (output_h - 1) % (input_h - 1)
                        and (output_w - 1) % (input_w - 1)):
                    warnings.warn(f'When align_corners={align_corners}, the output would more aligned if '
                                  f'input size {(input_h, input_w)} is `x+1` and out size {(output_h, output_w)} is `nx+1`')
    _lib.require_cuda(input, 'input')
    if input.dim() != 4:
        raise NotImplementedError('resize: only 4-D (N,C,H,W) inputs are on the segmentation path')
    if input.dtype not in _lib.LOGIT_DTYPES:
        raise TypeError('resize: float32, bfloat16 or float16 input expected, got %s' % input.dtype)
    H, W = _out_size(input, size, scale_factor)
    scales = _kernel_scales(scale_factor)
    if mode == 'bilinear':
        return _Resize.apply(input, H, W, 'bilinear', bool(align_corners), scales)
    if mode == 'nearest':
        if align_corners is not None:
            raise ValueError('align_corners option can only be set with the interpolating modes')
        return _Resize.apply(input, H, W, 'nearest', False, scales)
    raise NotImplementedError('resize: mode %r is not on the segmentation hot path (bilinear / nearest only)' % (mode,))


def add_prefix(inputs, prefix):
    """{name: v} -> {prefix.name: v} (utils/ops.py:29-45)."""
    return {f'{prefix}.{name}': value for name, value in inputs.items()}


class Upsample(nn.Module):
    """utils/ops.py:48-69."""

    def __init__(self, size=None, scale_factor=None, mode='nearest', align_corners=None):
        super().__init__()
        self.size = size
        if isinstance(scale_factor, tuple):
            self.scale_factor = tuple(float(factor) for factor in scale_factor)
        else:
            self.scale_factor = float(scale_factor) if scale_factor else None
        self.mode = mode
        self.align_corners = align_corners

    def forward(self, x):
        if not self.size:
            if isinstance(self.scale_factor, tuple):
                size = [int(t * f) for t, f in zip(x.shape[-2:], self.scale_factor)]
            else:
                size = [int(t * self.scale_factor) for t in x.shape[-2:]]
        else:
            size = self.size
        return resize(x, size, None, self.mode, self.align_corners)
