"""Fused entry for the decode-head call site (models/decode_heads/decode_head.py:261-321).

``fused_resize_losses`` does what ``BaseDecodeHead.losses`` does between the head's conv and the loss dict
— resize to the label size, every configured loss, top-1 accuracy — without materialising the up-sampled
logits and with one read of the logits per direction. ``B200DecodeHeadLossMixin`` packages it as a
``losses()`` method with the reference's signature for a head subclass registered under DECODEHEAD.
"""
import torch
import torch.nn as nn

from . import _lib
from .losses._function import LossSpec, run_fused
from .losses.accuracy import accuracy
from .losses.cross_entropy_loss import CrossEntropyLoss, _match_dtype
from .losses.dice_loss import DiceLoss
from .losses.tversky_loss import TverskyLoss
from .ops import resize


def _merge_spec(ce, dice, device, seg_weight, ignore_index, align_corners, want_acc):
    spec = LossSpec(align_corners=bool(align_corners), want_acc=want_acc, acc_ignore_index=ignore_index)
    if ce is not None:
        s = ce.spec(device, ignore_index=ignore_index)
        spec.want_ce = True
        spec.ce_reduction, spec.ce_class_weight, spec.ce_loss_weight = s.ce_reduction, s.ce_class_weight, s.ce_loss_weight
        spec.ce_ignore_index, spec.ce_avg_non_ignore, spec.single_pass = s.ce_ignore_index, s.ce_avg_non_ignore, s.single_pass
    if dice is not None:
        s = dice.spec(device)
        spec.want_dice = True
        spec.dice_reduction, spec.dice_class_weight, spec.dice_loss_weight = s.dice_reduction, s.dice_class_weight, s.dice_loss_weight
        spec.dice_ignore_index, spec.dice_smooth, spec.dice_exponent = s.dice_ignore_index, s.dice_smooth, s.dice_exponent
        spec.dice_mode, spec.tversky_alpha, spec.tversky_beta = s.dice_mode, s.tversky_alpha, s.tversky_beta
    return spec


def fused_resize_losses(seg_logit, seg_label, losses_decode, align_corners=False, ignore_index=255, seg_weight=None,
                        return_stats=False):
    """{loss_name: loss, ..., 'acc_seg': (1,) tensor} for low- or full-resolution ``seg_logit``.

    ``losses_decode`` is a loss module or a list / ModuleList of them (this package's CrossEntropyLoss and
    DiceLoss / TverskyLoss are fused into a single kernel launch; any other nn.Module loss is called on the materialised
    resize). Same-named losses are summed, as decode_head.py:283-293 does. ``return_stats`` adds the float64
    statistics vector of the fused launch under '_stats' (see distributed.py).
    """
    if isinstance(losses_decode, nn.Module) and not isinstance(losses_decode, nn.ModuleList):
        losses_decode = [losses_decode]
    losses_decode = list(losses_decode)
    H, W = int(seg_label.shape[-2]), int(seg_label.shape[-1])
    up = tuple(seg_logit.shape[2:]) != (H, W)

    fused_ce = fused_dice = None
    others = []
    for m in losses_decode:
        if type(m) is CrossEntropyLoss and not m.use_sigmoid and fused_ce is None and m.reduction != 'none':
            fused_ce = m
        elif type(m) is DiceLoss and fused_dice is None:
            fused_dice = m
        elif type(m) is TverskyLoss and fused_dice is None and (fused_ce is None or m.ignore_index == ignore_index):
            # Tversky masks its ignored pixels through the saved lse; it shares the launch with CE only when both ignore
            # the same label, otherwise it runs as its own (still fused forward/backward) call below
            fused_dice = m
        else:
            others.append(m)

    full = None
    if up:
        # The resize-fused single pass (csrc/loss_upgen.cuh) takes CE (+accuracy) for any up-sampling ratio with C <= 32.
        # Everything else is resized once (csrc/resize.cu: deterministic gather backward — there is no atomicAdd path):
        # dice needs per-class sums of the up-sampled soft-max, other modules expect label-resolution logits.
        n, c, h, w = (int(v) for v in seg_logit.shape)
        fast = (fused_dice is None and not others and fused_ce is not None and fused_ce.single_pass and seg_logit.is_cuda
                and _lib.load().b200seg_loss_fused_workspace_bytes(n, c, h, w, H, W, int(bool(align_corners))) > 0)
        if not fast:
            full = resize(seg_logit, size=(H, W), mode='bilinear', align_corners=align_corners, warning=False)
    src = full if full is not None else seg_logit

    out = {}

    def add(name, value):
        if name not in out:
            out[name] = value
        else:
            out[name] = out[name] + value

    # a lone sigmoid cross-entropy (the shipped default config): its launch also counts the top-1 hits, so the
    # accuracy needs no pass of its own over the logits
    if (fused_ce is None and fused_dice is None and len(others) == 1 and type(others[0]) is CrossEntropyLoss
            and others[0].use_sigmoid and not return_stats and src.is_cuda and src.dim() == 4 and src.shape[1] > 1):
        m = others[0]
        acc = torch.empty(1, dtype=torch.float32, device=src.device)
        label = seg_label.squeeze(1) if seg_label.dim() == 4 else seg_label
        out[m.loss_name] = m(src, label, weight=seg_weight, ignore_index=ignore_index, _acc_out=acc, _acc_ignore=ignore_index)
        out['acc_seg'] = acc
        return out

    spec = _merge_spec(fused_ce, fused_dice, seg_logit.device, seg_weight, ignore_index, align_corners, True)
    l_ce, l_dice, acc, log_vec = run_fused(src, seg_label, seg_weight, spec, with_log=True)
    # keep the reference's insertion order of the loss dict
    for m in losses_decode:
        if m is fused_ce:
            add(m.loss_name, _match_dtype(l_ce, seg_logit))
        elif m is fused_dice:
            add(m.loss_name, _match_dtype(l_dice, seg_logit))
        else:
            label = seg_label.squeeze(1) if seg_label.dim() == 4 else seg_label
            add(m.loss_name, m(src, label, weight=seg_weight, ignore_index=ignore_index))
    out['acc_seg'] = acc
    if return_stats:
        out['_stats'] = log_vec
    return out


class B200DecodeHeadLossMixin:
    """``losses()`` of the reference's BaseDecodeHead (decode_head.py:261-321) on the fused kernels.

    Use as ``class B200FCNHead(B200DecodeHeadLossMixin, FCNHead)`` and register it under DECODEHEAD; the
    head must expose ``loss_decode``, ``align_corners``, ``ignore_index`` and ``sampler`` as the reference does.

    Returned logits: with ``rescale=True`` (validation, utils/train_utils.py:115) the reference resizes the
    label-resolution logits again to the original image size(s) (:301-318); that is reproduced exactly,
    including the double interpolation. With ``rescale=False`` (training, train_utils.py:86, where the
    returned logits are discarded) the reference returns the label-resolution logits; materialising them
    would write the (N,C,H,W) tensor the fused loss avoids, so the head's own low-resolution logits are
    returned unless ``materialize_train_logits`` is set.
    """

    materialize_train_logits = False

    def losses(self, seg_logit, seg_label, meta_infos, rescale=False):
        assert isinstance(meta_infos, dict), 'the meta_infos in the losses function of the decode head must be a dict !'
        label_size = tuple(int(s) for s in seg_label.shape[2:])
        ori_img_size = meta_infos.get('ori_img_size_hw', None)
        sampler = getattr(self, 'sampler', None)
        need_full = (sampler is not None) or bool(rescale and ori_img_size) or self.materialize_train_logits
        full = None
        if need_full and tuple(seg_logit.shape[2:]) != label_size:
            full = resize(seg_logit, size=label_size, mode='bilinear', align_corners=self.align_corners)
        at_label = full if full is not None else seg_logit
        seg_weight = sampler.sample(at_label, seg_label) if sampler is not None else None
        loss = fused_resize_losses(at_label, seg_label, self.loss_decode, align_corners=self.align_corners,
                                   ignore_index=self.ignore_index, seg_weight=seg_weight)
        if rescale and ori_img_size:
            if isinstance(ori_img_size, tuple):
                rescaled = resize(at_label, size=ori_img_size, mode='bilinear', align_corners=self.align_corners)
            elif isinstance(ori_img_size, list):
                assert len(at_label) == len(ori_img_size)
                rescaled = [resize(at_label[i].unsqueeze(0), size=s, mode='bilinear', align_corners=self.align_corners)
                            for i, s in enumerate(ori_img_size)]
            else:
                rescaled = at_label
        else:
            rescaled = at_label
        return rescaled, loss
