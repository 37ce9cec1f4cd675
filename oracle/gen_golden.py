"""Generates tests/golden/hotpath_golden.npz by running the REFERENCE'S OWN FILES (oracle/ref_loader.py).

Run in the build container, where /root/reference exists:

    python -m oracle.gen_golden

The reference ships no golden vectors for this path, so these fixtures — seeded synthetic
inputs together with the outputs the unmodified reference produced for them on CPU (torch 2.11.0) — are
what pins the oracle and, through it, the CUDA path. Inputs are stored, not re-generated, so the fixtures
do not depend on the RNG.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

from . import ref_loader

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _labels(g, n, h, w, c, ignore=None, frac=0.1):
    y = torch.randint(0, c, (n, h, w), generator=g)
    if ignore is not None:
        m = torch.rand((n, h, w), generator=g) < frac
        y[m] = ignore
    return y


def loss_cases():
    """name -> dict(params) ; tensors are attached by build()."""
    C = []
    # ---- cross entropy at label resolution
    C.append(dict(name='ce_basic', kind='ce', shape=(2, 5, 8, 12), size=(8, 12), ignore=255, kw={}))
    C.append(dict(name='ce_sum_weighted', kind='ce', shape=(2, 5, 8, 12), size=(8, 12), ignore=255, pixel_weight=True,
                  kw=dict(class_weight=[0.5, 1.0, 1.5, 2.0, 0.25], reduction='sum', loss_weight=0.4)))
    C.append(dict(name='ce_none', kind='ce', shape=(2, 4, 6, 10), size=(6, 10), ignore=255, pixel_weight=True,
                  kw=dict(class_weight=[1.0, 2.0, 0.5, 1.5], reduction='none', loss_weight=2.0)))
    C.append(dict(name='ce_avg_non_ignore', kind='ce', shape=(3, 7, 9, 11), size=(9, 11), ignore=255,
                  kw=dict(avg_non_ignore=True, class_weight=[1, 2, 3, 4, 5, 6, 7.0])))
    C.append(dict(name='ce_avg_factor', kind='ce', shape=(2, 3, 8, 8), size=(8, 8), ignore=-100, avg_factor=37.0, kw={}))
    C.append(dict(name='ce_two_class', kind='ce', shape=(2, 2, 16, 16), size=(16, 16), ignore=-1, ignore_frac=0.0, kw={}))
    C.append(dict(name='ce_c40', kind='ce', shape=(1, 40, 6, 8), size=(6, 8), ignore=255, kw={}))
    # ---- resize fused in
    C.append(dict(name='ce_up4', kind='ce', shape=(2, 5, 4, 6), size=(16, 24), ignore=255, ac=False,
                  kw=dict(class_weight=[1, 2, 3, 4, 5.0])))
    C.append(dict(name='ce_up8', kind='ce', shape=(2, 19, 3, 5), size=(24, 40), ignore=255, ac=False, kw={}))
    C.append(dict(name='ce_up8_nonignore', kind='ce', shape=(1, 6, 4, 4), size=(32, 32), ignore=255, ac=False, pixel_weight=True,
                  kw=dict(avg_non_ignore=True)))
    C.append(dict(name='ce_up_ac1', kind='ce', shape=(2, 4, 5, 7), size=(17, 25), ignore=255, ac=True, kw={}))
    C.append(dict(name='ce_up_odd', kind='ce', shape=(1, 3, 5, 6), size=(13, 17), ignore=255, ac=False, kw={}))
    C.append(dict(name='ce_down', kind='ce', shape=(1, 3, 12, 10), size=(5, 4), ignore=255, ac=False, kw={}))
    # ---- dice
    C.append(dict(name='dice_basic', kind='dice', shape=(2, 4, 8, 8), size=(8, 8), ignore=255, kw={}))
    C.append(dict(name='dice_weighted', kind='dice', shape=(3, 5, 6, 10), size=(6, 10), ignore=255,
                  kw=dict(class_weight=[0.5, 1, 1.5, 2, 2.5], loss_weight=3.0, smooth=0.5)))
    C.append(dict(name='dice_ignore_in_range', kind='dice', shape=(2, 4, 8, 8), size=(8, 8), ignore=1, ignore_frac=0.0,
                  kw=dict(ignore_index=1)))
    C.append(dict(name='dice_exp3', kind='dice', shape=(2, 3, 8, 8), size=(8, 8), ignore=255, kw=dict(exponent=3)))
    C.append(dict(name='dice_c40', kind='dice', shape=(2, 40, 4, 8), size=(4, 8), ignore=255, kw={}))
    C.append(dict(name='dice_avg_factor', kind='dice', shape=(2, 3, 4, 4), size=(4, 4), ignore=255, avg_factor=3.0, kw={}))
    # ---- the decode-head chain: resize -> CE + Dice -> accuracy
    C.append(dict(name='head_ce_dice', kind='head', shape=(2, 6, 8, 8), size=(8, 8), ignore=255,
                  ce=dict(class_weight=[0.5, 0.7, 0.9, 1.1, 1.3, 1.5]), dice=dict(loss_weight=3.0)))
    C.append(dict(name='head_ce_dice_up', kind='head', shape=(2, 6, 4, 4), size=(16, 16), ignore=255, ac=False,
                  ce={}, dice=dict(loss_weight=3.0)))
    return C


def build(out_dir=OUT_DIR):
    ref = ref_loader.load()
    warnings.simplefilter('ignore')
    g = torch.Generator().manual_seed(20261018)
    data = {}
    manifest = {'torch': torch.__version__, 'cases': []}

    for case in loss_cases():
        n, c, h, w = case['shape']
        H, W = case['size']
        name = case['name']
        logits = (torch.randn((n, c, h, w), generator=g) * 2.0).requires_grad_(True)
        labels = _labels(g, n, H, W, c, case['ignore'] if case['ignore'] not in (None,) else None,
                         case.get('ignore_frac', 0.1))
        pw = torch.rand((n, H, W), generator=g) + 0.5 if case.get('pixel_weight') else None
        ac = case.get('ac', False)
        full = ref.resize(logits, size=(H, W), mode='bilinear', align_corners=ac, warning=False)
        outs = {}
        if case['kind'] == 'ce':
            mod = ref.CrossEntropyLoss(**case['kw'])
            loss = mod(full, labels, weight=pw, avg_factor=case.get('avg_factor'), ignore_index=case['ignore'])
            outs['loss'] = loss
            total = loss.sum() if loss.dim() else loss
        elif case['kind'] == 'dice':
            mod = ref.DiceLoss(**case['kw'])
            loss = mod(full, labels, avg_factor=case.get('avg_factor'), weight=pw, ignore_index=case['ignore'])
            outs['loss'] = loss
            total = loss
        else:
            ce = ref.CrossEntropyLoss(**case['ce'])
            dice = ref.DiceLoss(**case['dice'])
            l1 = ce(full, labels, weight=pw, ignore_index=case['ignore'])
            l2 = dice(full, labels, weight=pw, ignore_index=case['ignore'])
            outs['loss_ce'], outs['loss_dice'] = l1, l2
            total = l1 + l2
        if case['kind'] == 'ce' and case['kw'].get('reduction') == 'none':
            gsel = torch.rand(loss.shape, generator=g)  # a non-trivial upstream gradient for 'none'
            total = (loss * gsel).sum()
            data[name + '/grad_out'] = gsel.numpy()
        total.backward()
        outs['grad'] = logits.grad
        outs['acc'] = ref.accuracy(full.detach(), labels, ignore_index=case['ignore'] if case['ignore'] != -100 else None)
        data[name + '/logits'] = logits.detach().numpy()
        data[name + '/labels'] = labels.numpy()
        if pw is not None:
            data[name + '/pixel_weight'] = pw.numpy()
        for k, v in outs.items():
            data[name + '/' + k] = v.detach().numpy()
        manifest['cases'].append({k: v for k, v in case.items()})

    # ---- resize alone (utils/ops.py:7-26)
    for name, shape, size, ac in [('resize_up8', (1, 3, 4, 5), (32, 40), False), ('resize_ac1', (2, 2, 5, 7), (17, 25), True),
                                  ('resize_odd', (1, 2, 7, 5), (10, 16), False), ('resize_down', (1, 2, 12, 16), (5, 6), False)]:
        x = torch.randn(shape, generator=g, requires_grad=True)
        y = ref.resize(x, size=size, mode='bilinear', align_corners=ac, warning=False)
        go = torch.randn(y.shape, generator=g)
        y.backward(go)
        data[name + '/x'] = x.detach().numpy()
        data[name + '/y'] = y.detach().numpy()
        data[name + '/go'] = go.numpy()
        data[name + '/gx'] = x.grad.numpy()
        manifest['cases'].append(dict(name=name, kind='resize', shape=shape, size=size, ac=ac))
    x = torch.randn((1, 2, 4, 6), generator=g)
    data['resize_nearest/x'] = x.numpy()
    data['resize_nearest/y'] = ref.resize(x, size=(9, 15)).numpy()
    manifest['cases'].append(dict(name='resize_nearest', kind='resize_nearest', shape=(1, 2, 4, 6), size=(9, 15)))

    # ---- accuracy with top-k and threshold (models/losses/accuracy.py:6-61)
    # logits are made tie-free: torch.topk's tie order is unspecified
    p = torch.randn((64, 7), generator=g)
    t = torch.randint(0, 7, (64,), generator=g)
    data['acc_topk/pred'] = p.numpy()
    data['acc_topk/target'] = t.numpy()
    r = ref.accuracy(p, t, topk=(1, 3), thresh=0.2)
    data['acc_topk/out'] = np.stack([v.numpy() for v in r])
    p4 = torch.randn((2, 7, 6, 5), generator=g)
    t4 = _labels(g, 2, 6, 5, 7, 255)
    data['acc_topk4d/pred'] = p4.numpy()
    data['acc_topk4d/target'] = t4.numpy()
    r = ref.accuracy(p4, t4, topk=(1, 2, 5), ignore_index=255)
    data['acc_topk4d/out'] = np.stack([v.numpy() for v in r])
    manifest['cases'] += [dict(name='acc_topk', kind='acc', topk=(1, 3), thresh=0.2),
                          dict(name='acc_topk4d', kind='acc', topk=(1, 2, 5), ignore=255)]

    # ---- intersect_and_union (core/evaluation/metrics.py:210-270), incl. out-of-range values
    Cn, ign = 5, 255
    preds, gts = [], []
    for (hh, ww) in [(9, 13), (16, 16), (7, 5)]:
        pr = torch.randint(0, Cn, (hh, ww), generator=g)
        gt = torch.randint(0, Cn, (hh, ww), generator=g).float()
        gt[torch.rand((hh, ww), generator=g) < 0.15] = ign
        same = torch.rand((hh, ww), generator=g) < 0.5
        pr[same] = gt[same].long().clamp(0, Cn - 1)
        preds.append(pr)
        gts.append(gt)
    gts[0][0, 0] = 7.0     # out-of-range, not ignored: dropped from the label histogram, kept in pred
    preds[1][0, 1] = 9     # out-of-range prediction: dropped from pred
    gts[1][0, 1] = 9.0     # ... and pred == gt there: intersect value 9 is dropped as well
    a = ref_loader.intersect_and_union_cpu(ref, preds, gts, Cn, ign)
    for i, (pr, gt) in enumerate(zip(preds, gts)):
        data['iau/pred%d' % i] = pr.numpy()
        data['iau/gt%d' % i] = gt.numpy()
    data['iau/areas'] = np.stack([np.stack([x.numpy() for x in lst]) for lst in a], axis=1)  # (n,4,C) I,U,P,L
    manifest['cases'].append(dict(name='iau', kind='iau', num_classes=Cn, ignore=ign, n=3))

    # ---- process(): softmax -> argmax -> areas from logits. ignore_index = -1 as configs/dataset/KvasirSEG.py:8
    # (seg_metrics indexes class_names[ignore_index], metrics.py:201, so 255 raises IndexError in the reference)
    ign2 = -1
    gts2 = []
    for t in gts:
        t2 = t.clone()
        t2[t2 == ign] = ign2
        t2[t2 > Cn - 1] = 0
        gts2.append(t2)
    ev = ref.SegEvaluator(epoch=0, num_classes=Cn, class_names=['c%d' % i for i in range(Cn)], palette=None,
                          ignore_index=ign2, show_result=False)
    lg = [torch.randn((1, Cn, hh, ww), generator=g) * 3 for (hh, ww) in [(9, 13), (16, 16), (7, 5)]]
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a_, **k_: self
    try:
        ev.process(0, {'decode': [t.clone() for t in lg]}, {'ori_gt': [t.clone() for t in gts2]})
        met = ev.compute_metrics()
    finally:
        torch.Tensor.cuda = orig
    for i, t in enumerate(gts2):
        data['process/gt%d' % i] = t.numpy()
    for i, t in enumerate(lg):
        data['process/logits%d' % i] = t.numpy()
    data['process/areas'] = np.stack([np.stack([x.numpy() for x in lst]) for lst in ev.results['decode']], axis=1)
    for k in ('aAcc', 'mIoU', 'mAcc', 'mDice', 'mFscore', 'mPrecision', 'mRecall'):
        data['process/summary_' + k] = np.asarray(met['decode'][k])
    for k in ('IoU', 'Acc', 'Dice', 'Fscore', 'Precision', 'Recall'):
        data['process/class_' + k] = np.asarray(met['decode'][k])
    manifest['cases'].append(dict(name='process', kind='process', num_classes=Cn, ignore=ign2, n=3))

    # ---- total_area_to_metrics incl. 0/0 -> NaN, nan_to_num and beta (:272-356)
    I = torch.tensor([10., 0., 5., 0.]); P = torch.tensor([12., 0., 9., 4.]); L = torch.tensor([15., 0., 6., 0.])
    U = L + P - I
    for tag, kw in [('plain', {}), ('nan0_beta2', dict(nan_to_num=0, beta=2))]:
        r = ref.SegEvaluator.total_area_to_metrics(I, U, P, L, ['mIoU', 'mDice', 'mFscore'], **kw)
        for k, v in r.items():
            data['metrics_%s/%s' % (tag, k)] = np.asarray(v)
    data['metrics/I'], data['metrics/U'], data['metrics/P'], data['metrics/L'] = I.numpy(), U.numpy(), P.numpy(), L.numpy()
    manifest['cases'].append(dict(name='metrics', kind='metrics'))

    # ---- the reference's only known-answer example (models/losses/utils.py:95-111)
    l1 = ref.weighted_loss(lambda pred, target: (pred - target).abs())
    pr, tg, wt = torch.Tensor([0, 2, 3]), torch.Tensor([1, 1, 1]), torch.Tensor([1, 0, 1])
    data['kat/out'] = np.array([l1(pr, tg).item(), l1(pr, tg, wt).item(), l1(pr, tg, wt, avg_factor=2).item()], dtype=np.float32)
    data['kat/none'] = l1(pr, tg, reduction='none').numpy()

    # ---- sigmoid path: binary_cross_entropy + _expand_onehot_labels (cross_entropy_loss.py:77-164); appended last so
    # that the fixtures above do not change. class_weight (pos_weight) only on 2-D predictions: on (N,C,H,W) inputs
    # the reference's (C,) pos_weight broadcasts along W, not along the class dimension.
    bce_cases = [
        dict(name='bce_mean', shape=(2, 3, 8, 10), kw=dict()),
        dict(name='bce_sum_w', shape=(2, 4, 6, 6), pixel_weight=True, kw=dict(reduction='sum')),
        dict(name='bce_nonignore', shape=(2, 3, 8, 8), kw=dict(avg_non_ignore=True)),
        dict(name='bce_none', shape=(1, 3, 6, 8), pixel_weight=True, kw=dict(reduction='none')),
        dict(name='bce_single', shape=(2, 1, 8, 8), kw=dict()),
        dict(name='bce_2d_posw', shape=(40, 5), kw=dict(class_weight=[0.5, 1.0, 2.0, 1.5, 3.0])),
    ]
    for case in bce_cases:
        name, shape = case['name'], case['shape']
        x = (torch.randn(shape, generator=g) * 2).requires_grad_(True)
        if len(shape) == 4:
            ncls = max(shape[1], 2)
            y = torch.randint(0, ncls, (shape[0],) + shape[2:], generator=g)
            y[torch.rand(y.shape, generator=g) < 0.15] = 255
            w = torch.rand(y.shape, generator=g) + 0.5 if case.get('pixel_weight') else None
        else:
            y = torch.randint(0, shape[1], (shape[0],), generator=g)
            y[::7] = 255
            w = None
        mod = ref.CrossEntropyLoss(use_sigmoid=True, loss_weight=0.7, **case['kw'])
        loss = mod(x, y, weight=w, ignore_index=255)
        if loss.dim():
            go = torch.rand(loss.shape, generator=g)
            (loss * go).sum().backward()
            data[name + '/grad_out'] = go.numpy()
        else:
            loss.backward()
        data[name + '/logits'] = x.detach().numpy()
        data[name + '/labels'] = y.numpy()
        if w is not None:
            data[name + '/pixel_weight'] = w.numpy()
        data[name + '/loss'] = loss.detach().numpy()
        data[name + '/grad'] = x.grad.numpy()
        manifest['cases'].append(dict(name=name, kind='bce', shape=shape, kw=case['kw'], loss_weight=0.7, ignore=255,
                                      pixel_weight=bool(case.get('pixel_weight'))))

    # ---- TverskyLoss (models/losses/tversky_loss.py:24-148); appended last
    tv_cases = [
        dict(name='tversky_basic', shape=(2, 4, 8, 8), kw=dict()),
        dict(name='tversky_weighted', shape=(3, 5, 6, 10), kw=dict(class_weight=[0.5, 1, 1.5, 2, 2.5], loss_weight=2.0, smooth=0.5,
                                                                   alpha=0.4, beta=0.6)),
        dict(name='tversky_ignore_in_range', shape=(2, 4, 8, 8), kw=dict(ignore_index=1), ignore=1, ignore_frac=0.0),
        dict(name='tversky_c40', shape=(2, 40, 4, 8), kw=dict()),
    ]
    for case in tv_cases:
        name, shape = case['name'], case['shape']
        x = (torch.randn(shape, generator=g) * 2).requires_grad_(True)
        y = _labels(g, shape[0], shape[2], shape[3], shape[1], case.get('ignore', 255), case.get('ignore_frac', 0.1))
        loss = ref.TverskyLoss(**case['kw'])(x, y)
        loss.backward()
        data[name + '/logits'] = x.detach().numpy()
        data[name + '/labels'] = y.numpy()
        data[name + '/loss'] = loss.detach().numpy()
        data[name + '/grad'] = x.grad.numpy()
        manifest['cases'].append(dict(name=name, kind='tversky', shape=shape, kw=case['kw']))

    # ---- LovaszLoss (models/losses/lovasz_loss.py:26-312); appended last. Small segments: the reference's fp32 Jaccard
    # differences are accurate here (their absolute noise of ~6e-8 is far below increments of ~1/100), and no two errors
    # of a segment are closer than 1e-6 (checked below), so the sorted order — hence the gradient — is well defined.
    lov_cases = [
        dict(name='lovasz_present', shape=(2, 5, 8, 12), kw=dict(reduction='none')),
        dict(name='lovasz_all_weighted', shape=(2, 5, 8, 12), kw=dict(reduction='none', classes='all', loss_weight=2.0,
                                                                      class_weight=[0.5, 1.0, 1.5, 2.0, 0.25])),
        dict(name='lovasz_absent_class', shape=(2, 6, 8, 8), max_label=4, kw=dict(reduction='none')),
        dict(name='lovasz_all_absent_class', shape=(2, 6, 8, 8), max_label=4, kw=dict(reduction='none', classes='all')),
        dict(name='lovasz_list', shape=(2, 5, 8, 12), kw=dict(reduction='none', classes=[1, 3])),
        dict(name='lovasz_per_image_mean', shape=(3, 4, 8, 8), kw=dict(per_image=True, reduction='mean')),
        dict(name='lovasz_per_image_sum', shape=(3, 4, 8, 8), kw=dict(per_image=True, reduction='sum', class_weight=[1, 2, 3, 4.0])),
        dict(name='lovasz_per_image_none', shape=(3, 4, 8, 8), kw=dict(per_image=True, reduction='none')),
        dict(name='lovasz_per_image_avg', shape=(3, 4, 8, 8), avg_factor=2.5, kw=dict(per_image=True, reduction='mean')),
        dict(name='lovasz_c40', shape=(1, 40, 12, 16), kw=dict(reduction='none')),
        dict(name='lovasz_no_ignore', shape=(2, 3, 6, 10), ignore=None, kw=dict(reduction='none')),
        dict(name='lovasz_hinge', shape=(2, 1, 8, 12), kw=dict(loss_type='binary', reduction='none')),
        dict(name='lovasz_hinge_per_image', shape=(3, 1, 8, 8), kw=dict(loss_type='binary', per_image=True, reduction='mean',
                                                                        loss_weight=0.5)),
    ]
    def _lovasz_gap(x, y, shape, ign, binary, per_image):
        with torch.no_grad():
            if binary:
                err = [(1 - x[:, 0] * (2. * y.float() - 1))[y != 255]] if not per_image else \
                    [(1 - x[i, 0] * (2. * y[i].float() - 1))[y[i] != 255] for i in range(shape[0])]
            else:
                p = torch.softmax(x, 1)
                groups = [slice(i, i + 1) for i in range(shape[0])] if per_image else [slice(0, shape[0])]
                err = []
                for sl in groups:
                    keep = (y[sl] != ign) if ign is not None else torch.ones_like(y[sl], dtype=torch.bool)
                    for c in range(shape[1]):
                        err.append(((y[sl] == c).float() - p[sl, c]).abs()[keep])
            # margin of every adjacent pair of sorted errors over what two fp32 soft-max implementations can differ by
            # (3e-6 relative, plus two ulps of 1.0 for errors of the form 1 - p): > 1 means the order is unambiguous
            worst = float('inf')
            for e in err:
                if e.numel() > 1:
                    v = torch.sort(e.abs()).values
                    tol = 3e-6 * v[1:] + 2.4e-7 * (v[1:] > 0.5)
                    worst = min(worst, float(((v[1:] - v[:-1]) / tol).min()))
            return worst

    for case in lov_cases:
        name, shape = case['name'], case['shape']
        binary = case['kw'].get('loss_type') == 'binary'
        ign = case.get('ignore', 255)
        ncls = 2 if binary else case.get('max_label', shape[1])
        for _ in range(50):   # redraw until the sorted errors of every segment are separated (near-ties are common)
            x = (torch.randn(shape, generator=g) * 2).requires_grad_(True)
            y = _labels(g, shape[0], shape[2], shape[3], ncls, ign, 0.1)
            gap = _lovasz_gap(x, y, shape, ign, binary, bool(case['kw'].get('per_image')))
            if gap > 1.0:
                break
        mod = ref.LovaszLoss(**case['kw'])
        loss = mod(x, y, avg_factor=case.get('avg_factor'), ignore_index=ign)
        if loss.dim():
            go = torch.rand(loss.shape, generator=g)
            (loss * go).sum().backward()
            data[name + '/grad_out'] = go.numpy()
        else:
            loss.backward()
        assert gap > 1.0, (name, gap)
        data[name + '/logits'] = x.detach().numpy()
        data[name + '/labels'] = y.numpy()
        data[name + '/loss'] = loss.detach().numpy()
        data[name + '/grad'] = x.grad.numpy()
        manifest['cases'].append(dict(name=name, kind='lovasz', shape=shape, kw=case['kw'], ignore=ign,
                                      avg_factor=case.get('avg_factor'), order_margin=gap))

    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, 'hotpath_golden.npz'), **data)
    with open(os.path.join(out_dir, 'manifest.json'), 'w') as fh:
        json.dump(manifest, fh, indent=1, default=list)
    return data, manifest


if __name__ == '__main__':
    if not ref_loader.available():
        sys.exit('the reference tree is not available here; fixtures can only be generated in the build container')
    d, m = build()
    size = os.path.getsize(os.path.join(OUT_DIR, 'hotpath_golden.npz'))
    print('wrote %d arrays, %d cases, %d bytes' % (len(d), len(m['cases']), size))
