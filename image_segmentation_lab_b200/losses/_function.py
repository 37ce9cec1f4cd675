"""Host side of the fused logits -> loss path: one autograd.Function over the C ABI.

It plays the role of the call chain in the reference's ``BaseDecodeHead.losses``
(models/decode_heads/decode_head.py:261-295): resize -> CrossEntropyLoss / DiceLoss -> accuracy,
and of the ATen autograd graph behind it. All arithmetic happens in libb200seg.so; this file only
allocates buffers, fills descriptors and picks one of three execution plans:

  * ``up_single``   logits at lower resolution than the labels (any up-sampling ratio, both align_corners settings,
                    C <= 32), CE (+accuracy): forward+backward in one pass (b200seg_loss_fused_fwdbwd, combine
                    deferred to backward()).
  * ``flat_single`` logits at label resolution, CE (+accuracy), gradient needed: one pass.
  * ``two_pass``    everything else at label resolution (Dice, reduction='none', avg_non_ignore): b200seg_loss_fwd saves
                    the per-pixel log-sum-exp, b200seg_loss_bwd re-reads the logits once (CE + Dice with 32 < C <= 152:
                    the class-sliced TMA pipeline of csrc/loss_cs.cu, one read of the logits per direction).
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from .. import _lib

_EPS = float(torch.finfo(torch.float32).eps)


@dataclass
class LossSpec:
    align_corners: bool = False
    # cross entropy (models/losses/cross_entropy_loss.py:23-74)
    want_ce: bool = False
    ce_reduction: str = "mean"
    ce_class_weight: Optional[torch.Tensor] = None  # fp32 (C,) on the logits' device
    ce_loss_weight: float = 1.0
    ce_ignore_index: int = -100
    ce_avg_non_ignore: bool = False
    ce_avg_factor: Optional[float] = None
    # dice (models/losses/dice_loss.py:61-134)
    want_dice: bool = False
    dice_reduction: str = "mean"
    dice_class_weight: Optional[torch.Tensor] = None
    dice_loss_weight: float = 1.0
    dice_ignore_index: Optional[int] = 255
    dice_smooth: float = 1.0
    dice_exponent: float = 2.0
    dice_avg_factor: Optional[float] = None
    # tversky (models/losses/tversky_loss.py:24-148) shares the dice slot: masked sums, exponent 1
    dice_mode: str = "dice"
    tversky_alpha: float = 0.3
    tversky_beta: float = 0.7
    # accuracy (models/losses/accuracy.py:6-61, top-1)
    want_acc: bool = False
    acc_ignore_index: Optional[int] = None
    # execution
    single_pass: bool = True


def prep_labels(labels, logits):
    """(N,H,W) contiguous labels in a dtype the kernels read directly (no .long() pass)."""
    if labels.dim() == logits.dim() and labels.size(1) == 1:
        labels = labels.squeeze(1)
    if labels.dtype not in _lib.LABEL_DTYPES:
        labels = labels.long()
    if labels.device != logits.device:
        labels = labels.to(logits.device, non_blocking=True)
    return labels.contiguous()


def _none_int(v, default):
    return default if v is None else int(v)


class FusedLossFunction(torch.autograd.Function):
    """(logits, labels, pixel_weight, spec) -> (loss_ce, loss_dice, acc_seg, log_vec).

    ``log_vec`` is the float64 vector of additive per-call statistics (include/b200seg.h B200SEG_LOG_*): the
    payload of the single per-step all-reduce under data parallelism (distributed.py)."""

    @staticmethod
    def forward(ctx, logits, labels, pixel_weight, spec, grad_enabled=True):
        lib = _lib.load()
        _lib.require_cuda(logits, "logits")
        ctx.set_materialize_grads(False)   # unused outputs arrive as None in backward(), not as zero-filled tensors
        if logits.dtype not in _lib.LOGIT_DTYPES:
            raise TypeError("logits must be float32, bfloat16 or float16, got %s" % logits.dtype)
        if logits.dim() != 4:
            raise ValueError("logits must be (N,C,h,w), got shape %s" % (tuple(logits.shape),))
        logits_c = logits.contiguous()
        labels = prep_labels(labels, logits_c)
        if labels.dim() != 3 or labels.size(0) != logits_c.size(0):
            raise ValueError("labels must be (N,H,W) / (N,1,H,W) matching the logits batch, got %s" % (tuple(labels.shape),))
        N, Cc, h, w = logits_c.shape
        H, W = int(labels.shape[1]), int(labels.shape[2])
        dev = logits_c.device
        up = (h, w) != (H, W)
        if spec.want_dice and up:
            raise RuntimeError("dice needs logits at label resolution: resize first (fused_resize_losses does)")
        if spec.want_dice and not (spec.dice_exponent > 0):
            raise ValueError("DiceLoss exponent must be > 0")
        pw = None
        if pixel_weight is not None and spec.want_ce:
            pw = pixel_weight
            if pw.dim() == 4 and pw.size(1) == 1:
                pw = pw.squeeze(1)
            assert pw.dim() == 3 and tuple(pw.shape) == tuple(labels.shape), \
                "weight must have the shape of the per-pixel loss (models/losses/utils.py:62-64)"
            pw = pw.to(device=dev, dtype=torch.float32).contiguous()
        needs_grad = bool(ctx.needs_input_grad[0]) and bool(grad_enabled)   # needs_input_grad ignores no_grad()
        stream = _lib.stream_ptr(dev)

        with torch.cuda.device(dev):
            nc = N * Cc if spec.want_dice else 0
            ws = torch.empty(_lib.STATS_WORDS + _lib.LOG_WORDS + 4 * nc, dtype=torch.int64, device=dev)
            base = ws.data_ptr()
            stats_p = base
            log_p = base + 8 * _lib.STATS_WORDS
            part_p = log_p + 8 * _lib.LOG_WORDS
            coef_p = part_p + 24 * nc
            # The differentiable outputs are tensors of their own, NOT views of one workspace: the stock head does
            # `loss[name] += loss_decode(...)` for losses sharing a loss_name (decode_head.py:290), and autograd refuses
            # an in-place update of a view handed out by a multi-output Function.
            out_ce = torch.empty((), dtype=torch.float32, device=dev)
            out_dice = torch.empty((), dtype=torch.float32, device=dev)
            out_acc = torch.empty(1, dtype=torch.float32, device=dev)

            ce_none = spec.want_ce and spec.ce_reduction == "none"
            use_nvalid = bool(spec.want_ce and spec.ce_reduction == "mean" and spec.ce_avg_non_ignore
                              and spec.ce_avg_factor is None)
            # scale of d(loss_ce)/d(sum_px w*nll) known on the host (the n_valid case is resolved on device)
            ce_scale = float(spec.ce_loss_weight)
            if spec.want_ce and spec.ce_reduction == "mean":
                if spec.ce_avg_factor is not None:
                    ce_scale = ce_scale / float(torch.tensor(spec.ce_avg_factor + _EPS, dtype=torch.float32))
                elif not use_nvalid:
                    ce_scale = ce_scale / float(max(N * H * W, 1))

            fd = _lib.LossDesc()
            fd.logits = logits_c.data_ptr(); fd.labels = labels.data_ptr()
            fd.pixel_weight = pw.data_ptr() if pw is not None else None
            fd.ce_class_weight = spec.ce_class_weight.data_ptr() if (spec.want_ce and spec.ce_class_weight is not None) else None
            fd.logit_dtype = _lib.LOGIT_DTYPES[logits_c.dtype]; fd.label_dtype = _lib.LABEL_DTYPES[labels.dtype]
            fd.N, fd.C, fd.h, fd.w, fd.H, fd.W = N, Cc, h, w, H, W
            fd.align_corners = int(bool(spec.align_corners))
            fd.ignore_index = int(spec.ce_ignore_index)
            fd.acc_has_ignore = int(spec.acc_ignore_index is not None)
            fd.acc_ignore_index = _none_int(spec.acc_ignore_index, 0)
            fd.dice_ignore_index = _none_int(spec.dice_ignore_index, -(2 ** 62))
            tversky = spec.want_dice and spec.dice_mode == "tversky"
            stream_dice = spec.want_dice and (Cc > 32 or tversky)   # the streaming Dice kernels read lse back
            fd.dice_exponent = 1.0 if tversky else float(spec.dice_exponent)
            fd.dice_mode = _lib.MODE_TVERSKY if tversky else _lib.MODE_DICE
            fd.ce_loss_weight = float(spec.ce_loss_weight)
            fd.stats = stats_p
            fd.dice_part = part_p if spec.want_dice else None
            flags = (_lib.WANT_CE if spec.want_ce else 0) | (_lib.WANT_DICE if spec.want_dice else 0) | \
                    (_lib.WANT_ACC if spec.want_acc else 0)

            plan = "two_pass"
            if spec.single_pass and spec.want_ce and not spec.want_dice and not ce_none and N > 0:
                if up:
                    if lib.b200seg_loss_fused_workspace_bytes(N, Cc, h, w, H, W, fd.align_corners) > 0:
                        plan = "up_single"
                # float16 never takes this plan: its gradient is formed during the forward with the upstream gradient
                # still unknown, i.e. at loss_weight/(N*H*W) ~ 5e-7 * (p - onehot), below the float16 normal range — the
                # GradScaler's 65536 applied afterwards cannot bring the lost bits back. two_pass multiplies the scale
                # in fp32 before the single cast, as the reference's autograd does.
                elif needs_grad and not use_nvalid and logits_c.dtype != torch.float16 and lib.b200seg_loss_flat_single_ok(
                        logits_c.data_ptr(), labels.data_ptr(), fd.logit_dtype, fd.label_dtype, Cc, H * W, int(pw is not None)):
                    plan = "flat_single"   # bulk-copy pipeline (any C whose tile fits shared memory) or register tile (C <= 32)

            if up and plan == "two_pass" and needs_grad:
                # no scatter backward: shapes the resize-fused single pass does not take (C > 32, down-sampling, single_pass
                # off, reduction='none') are resized first — ops.resize has a deterministic gather backward
                raise RuntimeError("a gradient through low-resolution logits needs the resize-fused single pass (H >= h, W >= w, "
                                   "C <= 32, CE with 'mean' / 'sum' reduction); resize first (fused_resize_losses does)")
            loss_px = lse = grad = pb = None
            if plan == "two_pass":
                if ce_none:
                    loss_px = torch.empty((N, H, W), dtype=torch.float32, device=dev)
                    flags |= _lib.WANT_LOSS_PX
                    fd.loss_px = loss_px.data_ptr()
                if needs_grad or stream_dice:
                    lse = torch.empty((N, H, W), dtype=torch.float32, device=dev)
                    flags |= _lib.WANT_LSE
                    fd.lse = lse.data_ptr()
                fd.flags = flags
                _lib.check(lib.b200seg_loss_fwd(C.byref(fd), stream))
            else:
                fd.flags = flags
                fu = _lib.LossFusedDesc()
                fu.fwd = fd
                fu.grad_scale_host = ce_scale
                fu.use_nvalid = int(use_nvalid)
                if plan == "up_single":
                    if needs_grad:
                        nbytes = lib.b200seg_loss_fused_workspace_bytes(N, Cc, h, w, H, W, fd.align_corners)
                        pb = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
                        fu.workspace = pb.data_ptr()
                        fu.defer_combine = 1
                else:
                    grad = torch.empty_like(logits_c)
                    fu.grad_logits = grad.data_ptr()
                _lib.check(lib.b200seg_loss_fused_fwdbwd(C.byref(fu), stream))

            fin = _lib.FinalizeDesc()
            fin.stats = stats_p
            fin.dice_part = part_p if spec.want_dice else None
            fin.dice_class_weight = spec.dice_class_weight.data_ptr() if (spec.want_dice and spec.dice_class_weight is not None) else None
            fin.N, fin.C = N, Cc
            fin.n_pixels = N * H * W
            fin.ce_reduction = _lib.REDUCTIONS["sum" if ce_none else spec.ce_reduction]
            fin.ce_avg_non_ignore = int(spec.ce_avg_non_ignore)
            fin.ce_has_avg_factor = int(spec.ce_avg_factor is not None and not ce_none)
            fin.ce_avg_factor = float(spec.ce_avg_factor or 0.0)
            fin.dice_has_avg_factor = int(spec.dice_avg_factor is not None)
            fin.dice_avg_factor = float(spec.dice_avg_factor or 0.0)
            fin.ce_loss_weight = float(spec.ce_loss_weight)
            fin.dice_loss_weight = float(spec.dice_loss_weight)
            fin.dice_smooth = float(spec.dice_smooth)
            fin.dice_reduction = _lib.REDUCTIONS[spec.dice_reduction]
            fin.dice_ignore_index = fd.dice_ignore_index
            fin.dice_mode = fd.dice_mode
            fin.tversky_alpha = float(spec.tversky_alpha)
            fin.tversky_beta = float(spec.tversky_beta)
            fin.out_loss_ce = out_ce.data_ptr()
            fin.out_loss_dice = out_dice.data_ptr()
            fin.out_acc = out_acc.data_ptr()
            fin.dice_coef = coef_p if (spec.want_dice and needs_grad) else None
            fin.log_vec = log_p
            _lib.check(lib.b200seg_loss_finalize(C.byref(fin), stream))

        loss_ce = loss_px if ce_none else out_ce
        loss_dice = out_dice
        acc = out_acc
        log_vec = ws[_lib.STATS_WORDS:_lib.STATS_WORDS + _lib.LOG_WORDS].view(torch.float64)
        ctx.mark_non_differentiable(acc, log_vec)
        if needs_grad:
            ctx.spec = spec
            ctx.plan = plan
            ctx.ce_scale = ce_scale
            ctx.use_nvalid = use_nvalid
            ctx.ce_none = ce_none
            ctx.ws = ws
            ctx.geom = (N, Cc, h, w, H, W)
            ctx.ptrs = (stats_p, coef_p)
            ctx.consumed = False
            ctx.save_for_backward(logits_c, labels, pw if pw is not None else ws, lse if lse is not None else ws,
                                  grad if grad is not None else ws, pb if pb is not None else ws)
            ctx.has = (pw is not None, lse is not None, grad is not None, pb is not None)
        return loss_ce, loss_dice, acc, log_vec

    @staticmethod
    def backward(ctx, g_ce, g_dice, g_acc, g_log):
        lib = _lib.load()
        spec = ctx.spec
        logits, labels, pw, lse, grad, pb = ctx.saved_tensors
        has_pw, has_lse, has_grad, has_pb = ctx.has
        N, Cc, h, w, H, W = ctx.geom
        stats_p, coef_p = ctx.ptrs
        dev = logits.device
        stream = _lib.stream_ptr(dev)

        def scalar(g):
            if g is None:
                return torch.zeros((), dtype=torch.float32, device=dev)
            return g.detach().to(torch.float32).reshape(()).contiguous()

        with torch.cuda.device(dev):
            if ctx.plan == "up_single":
                gs = scalar(g_ce)
                out = torch.empty_like(logits)
                _lib.check(lib.b200seg_loss_fused_combine(
                    pb.data_ptr(), out.data_ptr(), _lib.LOGIT_DTYPES[logits.dtype], N, Cc, h, w,
                    C.c_float(ctx.ce_scale), gs.data_ptr(), int(ctx.use_nvalid), stats_p, stream))
                return out, None, None, None, None
            if ctx.plan == "flat_single":
                if ctx.consumed:
                    raise RuntimeError("the single-pass loss graph can be back-propagated once; build the loss with "
                                       "single_pass=False to call backward() repeatedly (retain_graph)")
                ctx.consumed = True
                gs = scalar(g_ce)
                _lib.check(lib.b200seg_scale_inplace(grad.data_ptr(), _lib.LOGIT_DTYPES[grad.dtype], grad.numel(),
                                                     gs.data_ptr(), stream))
                return grad, None, None, None, None

            bd = _lib.LossBwdDesc()
            bd.logits = logits.data_ptr(); bd.labels = labels.data_ptr()
            bd.pixel_weight = pw.data_ptr() if has_pw else None
            bd.ce_class_weight = spec.ce_class_weight.data_ptr() if (spec.want_ce and spec.ce_class_weight is not None) else None
            bd.lse = lse.data_ptr()
            bd.logit_dtype = _lib.LOGIT_DTYPES[logits.dtype]; bd.label_dtype = _lib.LABEL_DTYPES[labels.dtype]
            bd.N, bd.C, bd.h, bd.w, bd.H, bd.W = N, Cc, h, w, H, W
            bd.align_corners = int(bool(spec.align_corners))
            bd.ignore_index = int(spec.ce_ignore_index)
            bd.dice_ignore_index = _none_int(spec.dice_ignore_index, -(2 ** 62))
            tversky = spec.want_dice and spec.dice_mode == "tversky"
            bd.dice_exponent = 1.0 if tversky else float(spec.dice_exponent)
            bd.dice_mode = _lib.MODE_TVERSKY if tversky else _lib.MODE_DICE
            bd.stats = stats_p
            keep = []
            want_ce = spec.want_ce and g_ce is not None
            want_dice = spec.want_dice and g_dice is not None
            if spec.want_dice and not want_dice:
                g_dice = torch.zeros((), dtype=torch.float32, device=dev)
                want_dice = True
            bd.flags = (_lib.WANT_CE if want_ce else 0) | (_lib.WANT_DICE if want_dice else 0)
            out = torch.empty_like(logits)
            if not (want_ce or want_dice):
                return out.zero_(), None, None, None, None
            if want_ce:
                bd.ce_use_nvalid = int(ctx.use_nvalid)
                if ctx.ce_none:
                    gpx = g_ce.detach().to(torch.float32).expand(N, H, W).contiguous()
                    keep.append(gpx)
                    bd.ce_grad_px = gpx.data_ptr()
                    bd.ce_scale_host = float(spec.ce_loss_weight)
                else:
                    gs = scalar(g_ce)
                    keep.append(gs)
                    bd.ce_grad_out = gs.data_ptr()
                    bd.ce_scale_host = ctx.ce_scale
            if want_dice:
                gd = scalar(g_dice)
                keep.append(gd)
                bd.dice_grad_out = gd.data_ptr()
                bd.dice_coef = coef_p
            bd.grad_logits = out.data_ptr()
            if want_dice and (Cc > 32 or tversky):
                dot = torch.empty((N, H, W), dtype=torch.float32, device=dev)
                keep.append(dot)
                bd.scratch_px = dot.data_ptr()
            _lib.check(lib.b200seg_loss_bwd(C.byref(bd), stream))
        return out, None, None, None, None


def run_fused(logits, labels, pixel_weight, spec, with_log=False):
    loss_ce, loss_dice, acc, log_vec = FusedLossFunction.apply(logits, labels, pixel_weight, spec, torch.is_grad_enabled())
    if with_log:
        return loss_ce, loss_dice, acc, log_vec
    return loss_ce, loss_dice, acc
