"""GPU: parity of the CUDA path (through the Python mirror of the reference interface, which calls the C ABI)
against (1) the fixtures recorded from the reference's own files and (2) the oracle on the same seeded inputs.

Gates: areas bit-exact; fp32 loss <= 1e-5 relative, fp32 gradients <= 1e-4 relative
(inf-norm over inf-norm); bf16 / fp16 logits <= 2**-7 relative against the oracle on the fp32-upcast inputs.
"""
import warnings

import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.helpers import (case_tensors, loss_case_cuda, loss_case_oracle, rel_err, synth_labels, synth_logits)

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-5
GRAD_TOL = 1e-4
HALF_TOL = 2.0 ** -7


@pytest.fixture(scope='module')
def B():
    import image_segmentation_lab_b200 as pkg
    pkg.load_library()
    warnings.simplefilter('ignore')
    return pkg


def _loss_cases(manifest):
    return [c for c in manifest['cases'] if c['kind'] in ('ce', 'dice', 'head')]


def _check(out, ref, name, loss_tol=LOSS_TOL, grad_tol=GRAD_TOL, acc_tol=1e-3, loss_atol=0.0):
    for k in ref:
        if k.startswith('loss') and loss_atol and float((out[k].double().cpu() - ref[k].double().cpu()).abs().max()) <= loss_atol:
            continue   # cancellation-dominated value (1 - num/den ~ 1e-3): an fp32 ulp of the terms is the floor
        if k == 'grad':
            assert rel_err(out[k], ref[k]) <= grad_tol, '%s grad rel err %.3e' % (name, rel_err(out[k], ref[k]))
        elif k == 'acc':
            assert abs(float(out[k]) - float(ref[k])) <= acc_tol, '%s acc %r vs %r' % (name, float(out[k]), float(ref[k]))
        else:
            assert rel_err(out[k], ref[k]) <= loss_tol, '%s %s rel err %.3e' % (name, k, rel_err(out[k], ref[k]))


# ------------------------------------------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize('single_pass', [True, False])
def test_golden_loss_cases_fp32(B, golden, single_pass):
    data, manifest = golden
    for case in _loss_cases(manifest):
        name = case['name']
        logits, labels, pw = case_tensors(data, name, device='cuda')
        go = torch.from_numpy(data[name + '/grad_out']).cuda() if (name + '/grad_out') in data else None
        out = loss_case_cuda(case, logits, labels, pw, go, single_pass=single_pass)
        ref = {k: torch.from_numpy(data[name + '/' + k]) for k in out}
        _check(out, ref, name + ('' if single_pass else '[two_pass]'))


@pytest.mark.parametrize('label_dtype', [torch.uint8, torch.int32, torch.float32, torch.float64, torch.int16])
def test_golden_loss_cases_label_dtypes(B, golden, label_dtype):
    """Labels are consumed in the dtype the data pipeline delivers."""
    data, manifest = golden
    ran = 0
    for case in _loss_cases(manifest):
        if case['ignore'] < 0 and label_dtype == torch.uint8:
            continue
        name = case['name']
        logits, labels, pw = case_tensors(data, name, device='cuda')
        go = torch.from_numpy(data[name + '/grad_out']).cuda() if (name + '/grad_out') in data else None
        out = loss_case_cuda(case, logits, labels, pw, go, label_dtype=label_dtype)
        ref = {k: torch.from_numpy(data[name + '/' + k]) for k in out}
        _check(out, ref, '%s[%s]' % (name, label_dtype))
        ran += 1
    assert ran >= 15


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
def test_golden_loss_cases_half(B, golden, dtype):
    """16-bit logits: fp32 math inside; compared with the oracle on the fp32-upcast of the SAME rounded logits."""
    data, manifest = golden
    for case in _loss_cases(manifest):
        name = case['name']
        logits, labels, pw = case_tensors(data, name, device='cuda')
        lq = logits.to(dtype)
        go = torch.from_numpy(data[name + '/grad_out']).cuda() if (name + '/grad_out') in data else None
        out = loss_case_cuda(case, lq, labels, pw, go)
        ref = loss_case_oracle(case, lq.float(), labels, pw, go)
        assert out['grad'].dtype == dtype
        _check(out, ref, '%s[%s]' % (name, dtype), loss_tol=HALF_TOL, grad_tol=2 * HALF_TOL, acc_tol=0.5)


def test_golden_resize(B, golden):
    data, manifest = golden
    for case in [c for c in manifest['cases'] if c['kind'] == 'resize']:
        name = case['name']
        x = torch.from_numpy(data[name + '/x']).cuda().requires_grad_(True)
        y = B.resize(x, size=tuple(case['size']), mode='bilinear', align_corners=case['ac'], warning=False)
        y.backward(torch.from_numpy(data[name + '/go']).cuda())
        assert rel_err(y, data[name + '/y']) <= 2e-6, name     # the fixtures were recorded with ATen's CPU kernel
        assert rel_err(x.grad, data[name + '/gx']) <= 1e-5, name
        # against ATen's CUDA kernel on this GPU the forward is bit-identical (FMA contraction pinned, csrc/common.cuh)
        assert torch.equal(y, torch.nn.functional.interpolate(x, size=tuple(case['size']), mode='bilinear', align_corners=case['ac'])), name
    y = B.resize(torch.from_numpy(data['resize_nearest/x']).cuda(), size=(9, 15))
    np.testing.assert_array_equal(y.cpu().numpy(), data['resize_nearest/y'])
    x = torch.randn(2, 3, 5, 7, device='cuda')
    assert torch.equal(B.resize(x, size=(5, 7), mode='bilinear', align_corners=False), x)   # same size: a copy
    up = B.Upsample(scale_factor=2, mode='bilinear', align_corners=False)
    assert torch.equal(up(x), torch.nn.functional.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False))
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        for shp, size, ac in (((2, 5, 37, 53), (111, 160), True), ((1, 7, 65, 129), (513, 1025), False), ((3, 2, 40, 30), (17, 19), False)):
            xx = (torch.randn(shp, device='cuda') * 3).to(dt)
            assert torch.equal(B.resize(xx, size=size, mode='bilinear', align_corners=ac, warning=False),
                               torch.nn.functional.interpolate(xx, size=size, mode='bilinear', align_corners=ac)), (dt, shp, size, ac)


def test_golden_accuracy_topk(B, golden):
    data, _ = golden
    r = B.accuracy(torch.from_numpy(data['acc_topk/pred']).cuda(), torch.from_numpy(data['acc_topk/target']).cuda(),
                   topk=(1, 3), thresh=0.2)
    np.testing.assert_allclose(np.stack([v.cpu().numpy() for v in r]), data['acc_topk/out'], rtol=1e-6)
    r = B.accuracy(torch.from_numpy(data['acc_topk4d/pred']).cuda(), torch.from_numpy(data['acc_topk4d/target']).cuda(),
                   topk=(1, 2, 5), ignore_index=255)
    np.testing.assert_allclose(np.stack([v.cpu().numpy() for v in r]), data['acc_topk4d/out'], rtol=1e-6)
    single = B.accuracy(torch.from_numpy(data['acc_topk4d/pred']).cuda(), torch.from_numpy(data['acc_topk4d/target']).cuda(),
                        ignore_index=255)
    assert single.shape == (1,) and abs(float(single) - float(data['acc_topk4d/out'][0])) < 1e-4
    mod = B.Accuracy(topk=(1, 2, 5), ignore_index=255)
    r2 = mod(torch.from_numpy(data['acc_topk4d/pred']).cuda(), torch.from_numpy(data['acc_topk4d/target']).cuda())
    np.testing.assert_allclose(np.stack([v.cpu().numpy() for v in r2]), data['acc_topk4d/out'], rtol=1e-6)
    empty = B.accuracy(torch.zeros(0, 7, device='cuda'), torch.zeros(0, dtype=torch.long, device='cuda'))
    assert float(empty) == 0.0


def test_golden_intersect_and_union_bit_exact(B, golden):
    data, _ = golden
    preds = [torch.from_numpy(data['iau/pred%d' % i]).cuda() for i in range(3)]
    gts = [torch.from_numpy(data['iau/gt%d' % i]) for i in range(3)]      # CPU float32, as the data loader delivers
    lists = B.SegEvaluator.intersect_and_union(preds, gts, 5, 255)
    assert len(lists) == 4 and all(len(l) == 3 for l in lists)
    for j in range(4):
        got = np.stack([t.numpy() for t in lists[j]])
        assert got.dtype == np.float32 and not lists[j][0].is_cuda
        np.testing.assert_array_equal(got, data['iau/areas'][:, j])
    areas = B.areas_device(preds, [g.cuda() for g in gts], 5, 255)
    assert areas.dtype == torch.int64 and areas.is_cuda
    np.testing.assert_array_equal(areas.cpu().numpy(), data['iau/areas'][:, [0, 2, 3]].astype(np.int64))
    for pdt, gdt in [(torch.int32, torch.int64), (torch.uint8, torch.uint8), (torch.int64, torch.float64)]:
        a2 = B.areas_device([p.clamp(0, 255).to(pdt) for p in preds], [g.to(gdt).cuda() for g in gts], 5, 255)
        np.testing.assert_array_equal(a2.cpu().numpy(), data['iau/areas'][:, [0, 2, 3]].astype(np.int64))


def test_golden_process_and_metrics(B, golden, capsys):
    data, _ = golden
    logits = [torch.from_numpy(data['process/logits%d' % i]).cuda() for i in range(3)]
    gts = [torch.from_numpy(data['process/gt%d' % i]) for i in range(3)]
    ev = B.SegEvaluator(epoch=0, num_classes=5, class_names=['c%d' % i for i in range(5)], palette=None, ignore_index=-1,
                        show_result=False)
    pred_batch = {'decode': [t.clone() for t in logits], 'aux': [t.clone() for t in logits]}
    ev.process(0, pred_batch, {'ori_gt': gts})
    ev.process(1, {'decode': [t.clone() for t in logits[:1]]}, {'ori_gt': gts[:1]})
    res = ev.results
    assert set(res.keys()) == {'decode', 'aux'} and len(res['decode'][0]) == 4 and len(res['aux'][0]) == 3
    got = np.stack([np.stack([t.numpy() for t in res['aux'][j]]) for j in range(4)], axis=1)
    np.testing.assert_array_equal(got, data['process/areas'])
    ev2 = B.SegEvaluator(epoch=0, num_classes=5, class_names=['c%d' % i for i in range(5)], palette=None, ignore_index=-1,
                         show_result=False, keep_pred_maps=True)
    pb = {'decode': [t.clone() for t in logits]}
    ev2.process(0, pb, {'ori_gt': gts})
    for i in range(3):   # the reference replaces the logits by label maps in place (metrics.py:107)
        assert torch.equal(pb['decode'][i].cpu(), O.argmax_labels(logits[i].cpu()))
    met = ev2.compute_metrics()['decode']
    for k in ('aAcc', 'mIoU', 'mAcc', 'mDice', 'mFscore', 'mPrecision', 'mRecall'):
        assert met[k] == data['process/summary_' + k], k
    for k in ('IoU', 'Acc', 'Dice', 'Fscore', 'Precision', 'Recall'):
        np.testing.assert_array_equal(met[k], data['process/class_' + k])
    assert torch.equal(ev2.area_totals('decode'), torch.from_numpy(data['process/areas'].astype(np.int64).sum(0)))


def test_golden_sigmoid_bce(B, golden):
    """CrossEntropyLoss(use_sigmoid=True): the shipped default config (configs/network/deeplabv3/*.py:30,41)."""
    data, manifest = golden
    for case in [c for c in manifest['cases'] if c['kind'] == 'bce']:
        name, kw = case['name'], dict(case['kw'])
        for dtype, ltol, gtol in ((torch.float32, LOSS_TOL, GRAD_TOL), (torch.bfloat16, HALF_TOL, 2 * HALF_TOL)):
            x = torch.from_numpy(data[name + '/logits']).cuda().to(dtype).requires_grad_(True)
            y = torch.from_numpy(data[name + '/labels']).cuda()
            w = torch.from_numpy(data[name + '/pixel_weight']).cuda() if case['pixel_weight'] else None
            mod = B.CrossEntropyLoss(use_sigmoid=True, loss_weight=case['loss_weight'], **kw)
            loss = mod(x, y, weight=w, ignore_index=255)
            go = torch.from_numpy(data[name + '/grad_out']).cuda() if loss.dim() else None
            ((loss.float() * go).sum() if go is not None else loss).backward()
            if dtype == torch.float32:
                ref_loss, ref_grad = data[name + '/loss'], data[name + '/grad']
            else:   # the oracle on the fp32 upcast of the same rounded logits
                xo = x.detach().float().requires_grad_(True)
                kwo = dict(kw)
                if 'class_weight' in kwo:
                    kwo['class_weight'] = xo.new_tensor(kwo['class_weight'])
                lo = case['loss_weight'] * O.binary_cross_entropy(xo, y, w, ignore_index=255, **kwo)
                ((lo * go).sum() if go is not None else lo).backward()
                ref_loss, ref_grad = lo.detach(), xo.grad
            assert loss.shape == tuple(np.shape(ref_loss)) and loss.dtype == dtype
            assert rel_err(loss, ref_loss) <= ltol, '%s[%s] loss %.3e' % (name, dtype, rel_err(loss, ref_loss))
            assert rel_err(x.grad, ref_grad) <= gtol, '%s[%s] grad %.3e' % (name, dtype, rel_err(x.grad, ref_grad))
    # full Kvasir-like shape: 2 classes, 256x256 (BASELINE config 1 with the sigmoid loss), against the oracle on GPU
    x = synth_logits((2, 2, 256, 256), 21, device='cuda').requires_grad_(True)
    y = synth_labels((2, 256, 256), 2, 21, ignore_index=255, device='cuda')
    loss = B.CrossEntropyLoss(use_sigmoid=True)(x, y, ignore_index=255)
    loss.backward()
    xo = x.detach().clone().requires_grad_(True)
    lo = O.binary_cross_entropy(xo, y, ignore_index=255)
    lo.backward()
    assert rel_err(loss, lo) <= LOSS_TOL and rel_err(x.grad, xo.grad) <= GRAD_TOL


# ------------------------------------------------------------------------------------------------ oracle, larger shapes
def _head_case(B, shape, size, C, dtype, ce_kw, dice_kw, ac=False, ignore=255, seed=0, pixel_weight=False,
               loss_tol=LOSS_TOL, grad_tol=GRAD_TOL, single_pass=True, margin=True, loss_atol=0.0, acc_tol=None):
    n = shape[0]
    logits = synth_logits(shape, seed, dtype=dtype, device='cuda', margin=margin)
    labels = synth_labels((n,) + tuple(size), C, seed, ignore_index=ignore, block=8, device='cuda')
    pw = (torch.rand((n,) + tuple(size), device='cuda') + 0.5) if pixel_weight else None
    case = dict(kind='head' if dice_kw is not None else 'ce', size=size, ac=ac, ignore=ignore)
    if dice_kw is not None:
        case['ce'], case['dice'] = ce_kw, dice_kw
    else:
        case['kw'] = ce_kw
    out = loss_case_cuda(case, logits, labels, pw, single_pass=single_pass)
    ref = loss_case_oracle(case, logits.float(), labels, pw)   # the unfused ATen chain on the same GPU
    if acc_tol is None:
        acc_tol = 1e-3 if dtype == torch.float32 else 0.5
    _check(out, ref, 'head%s' % (shape,), loss_tol=loss_tol, grad_tol=grad_tol, acc_tol=acc_tol, loss_atol=loss_atol)
    return out, ref


def test_config1_unet_shape(B):
    """BASELINE config 1: 2x2x256x256, 2 classes, CE, no ignore (ignore_index=-1 as configs/dataset/KvasirSEG.py:8)."""
    _head_case(B, (2, 2, 256, 256), (256, 256), 2, torch.float32, {}, None, ignore=-1)
    _head_case(B, (2, 2, 256, 256), (256, 256), 2, torch.float32, {}, None, ignore=-1, single_pass=False)


@pytest.mark.parametrize('ac', [False, True])
def test_config2_cityscapes_shape(B, ac):
    """BASELINE config 2 at full size: (8,19,64,128) -> 512x1024, CE + ignore_index=255, both align_corners."""
    out, ref = _head_case(B, (8, 19, 64, 128), (512, 1024), 19, torch.float32, {}, None, ac=ac, acc_tol=0.01)
    assert out['grad'].shape == (8, 19, 64, 128)


def test_config2_variants(B):
    # The thread-per-cell kernel decides top-1 from the exponentials it already holds (label counted when its term is
    # within 2e-6 of the row maximum), so a pixel whose two best interpolated logits are closer than that can flip
    # against torch.topk (about 10 of a million pixels here): the logged accuracy is allowed 0.01 %
    t = 0.01
    _head_case(B, (2, 19, 32, 64), (512, 1024), 19, torch.float32, dict(avg_non_ignore=True), None, pixel_weight=True, acc_tol=t)  # S=16
    _head_case(B, (2, 19, 128, 256), (512, 1024), 19, torch.float32, dict(class_weight=[1.0 + 0.05 * i for i in range(19)]), None,
               acc_tol=t)  # S=4
    _head_case(B, (1, 21, 16, 16), (512, 512), 21, torch.float32, dict(reduction='sum'), None, acc_tol=t)  # S=32
    _head_case(B, (2, 32, 40, 24), (320, 192), 32, torch.float32, {}, None, acc_tol=t)  # C=32, non power-of-two extents
    # nx+1 sizes, align_corners=True: the thread-per-cell kernel. At a non-dyadic ratio its interpolation weights differ
    # from ATen's in the last bits, which flips the top-1 of pixels whose two best classes are closer than ~1e-5 (about 20
    # of a million here): the logged accuracy is allowed 0.01 %
    _head_case(B, (2, 19, 65, 129), (513, 1025), 19, torch.float32, {}, None, ac=True, acc_tol=0.01)
    _head_case(B, (2, 150, 32, 32), (256, 256), 150, torch.float32, {}, None)  # C > 32 -> general path
    _head_case(B, (2, 19, 64, 128), (512, 1024), 19, torch.bfloat16, {}, None, loss_tol=HALF_TOL, grad_tol=2 * HALF_TOL)



def test_resize_fused_extreme_logits_and_small_shapes(B):
    """Resize-fused single pass: steep logits (a pixel's classes all far below its row's upper bound -> the exact
    per-pixel-max redo path), every label dtype's edge handling on tiny and odd low-res extents, S in {4, 8, 16, 32}."""
    for S, (h, w), C, n in ((8, (5, 7), 19, 2), (4, (3, 9), 5, 3), (16, (2, 3), 21, 2), (32, (1, 2), 32, 1), (8, (1, 1), 3, 2),
                             (8, (9, 4), 2, 2), (4, (6, 6), 13, 1)):
        for scale in (1.0, 60.0):
            g = torch.Generator().manual_seed(S * 100 + C)
            x = (torch.randn((n, C, h, w), generator=g) * 3.0 * scale).cuda()
            labels = synth_labels((n, h * S, w * S), C, S + C, ignore_index=255, block=3, device='cuda')
            case = dict(kind='ce', size=(h * S, w * S), ac=False, ignore=255, kw=dict(class_weight=[0.5 + 0.1 * i for i in range(C)]))
            out = loss_case_cuda(case, x, labels, None)
            ref = loss_case_oracle(case, x, labels, None)
            _check(out, ref, 'S%d C%d %dx%d x%g' % (S, C, h, w, scale), loss_tol=2e-5 if scale > 1 else LOSS_TOL, acc_tol=0.05,
                   loss_atol=0.0)



def test_tversky_golden_and_oracle(B, golden):
    """TverskyLoss (models/losses/tversky_loss.py:24-148): the reference's fixtures in fp32, the oracle on larger shapes
    (C below and above 32, bf16), and the launch shared with CrossEntropyLoss through fused_resize_losses."""
    data, manifest = golden
    cases = [c for c in manifest['cases'] if c['kind'] == 'tversky']
    assert len(cases) >= 4
    for case in cases:
        name = case['name']
        x = torch.from_numpy(data[name + '/logits']).cuda().requires_grad_(True)
        y = torch.from_numpy(data[name + '/labels']).cuda()
        loss = B.TverskyLoss(**case['kw'])(x, y, weight=None, ignore_index=255)   # call-site kwargs are swallowed
        loss.backward()
        assert rel_err(loss, data[name + '/loss']) <= LOSS_TOL, '%s loss %.3e' % (name, rel_err(loss, data[name + '/loss']))
        assert rel_err(x.grad, data[name + '/grad']) <= GRAD_TOL, '%s grad %.3e' % (name, rel_err(x.grad, data[name + '/grad']))
    for shape, C, dtype, kw in (((4, 19, 96, 160), 19, torch.float32, dict(alpha=0.4, beta=0.6, smooth=0.5)),
                                ((2, 150, 64, 64), 150, torch.float32, dict(class_weight=torch.linspace(0.5, 1.5, 150).tolist())),
                                ((2, 21, 128, 128), 21, torch.bfloat16, dict(loss_weight=2.0)),
                                ((3, 7, 37, 53), 7, torch.float32, dict(ignore_index=3))):   # odd extent, in-range ignore
        x = synth_logits(shape, 31, dtype=dtype, device='cuda').requires_grad_(True)
        y = synth_labels((shape[0],) + shape[2:], C, 31, ignore_index=255, block=8, device='cuda')
        loss = B.TverskyLoss(**kw)(x, y)
        loss.backward()
        xo = x.detach().float().requires_grad_(True)
        lo = O.tversky_loss_module(xo, y, **kw)
        lo.backward()
        tol = (LOSS_TOL, GRAD_TOL) if dtype == torch.float32 else (HALF_TOL, 2 * HALF_TOL)
        assert rel_err(loss, lo) <= tol[0], '%s loss %.3e' % (shape, rel_err(loss, lo))
        assert rel_err(x.grad, xo.grad) <= tol[1], '%s grad %.3e' % (shape, rel_err(x.grad, xo.grad))
    # decode-head call: CE + Tversky on low-resolution logits, one fused launch set
    x = synth_logits((2, 19, 32, 64), 5, device='cuda').requires_grad_(True)
    y = synth_labels((2, 128, 256), 19, 5, ignore_index=255, block=8, device='cuda').unsqueeze(1)
    out = B.fused_resize_losses(x, y, [B.CrossEntropyLoss(), B.TverskyLoss(loss_weight=0.5)], ignore_index=255)
    assert list(out.keys()) == ['loss_ce', 'loss_tversky', 'acc_seg']
    (out['loss_ce'] + out['loss_tversky']).backward()
    xo = x.detach().clone().requires_grad_(True)
    full = O.resize(xo, size=(128, 256), mode='bilinear', align_corners=False)
    ref_ce = O.cross_entropy_loss_module(full, y.squeeze(1), ignore_index=255)
    ref_tv = O.tversky_loss_module(full, y.squeeze(1), loss_weight=0.5)
    (ref_ce + ref_tv).backward()
    assert rel_err(out['loss_ce'], ref_ce) <= LOSS_TOL and rel_err(out['loss_tversky'], ref_tv) <= LOSS_TOL
    assert rel_err(x.grad, xo.grad) <= GRAD_TOL



def test_bulk_pipeline_single_pass_shapes(B):
    """Label-resolution single-pass CE (cp.async.bulk pipeline, csrc/loss_bulk.cu): many classes, 16-bit logits, partial
    last tiles (H*W not a multiple of the 256-pixel tile), uint8 / int32 labels, class weights, more tiles than CTAs."""
    cw150 = torch.linspace(0.5, 1.5, 150).tolist()
    for shape, C, dtype, ldt, kw in (((2, 150, 64, 96), 150, torch.float32, torch.int64, dict(class_weight=cw150)),
                                     ((2, 150, 48, 40), 150, torch.bfloat16, torch.uint8, {}),
                                     ((3, 33, 20, 12), 33, torch.float32, torch.int32, {}),        # 240 px: one partial tile
                                     ((1, 21, 36, 52), 21, torch.float16, torch.int64, {}),        # 1872 px = 7 tiles + 80 px
                                     ((4, 100, 160, 160), 100, torch.float32, torch.uint8, dict(loss_weight=0.4)),
                                     ((2, 260, 32, 32), 260, torch.bfloat16, torch.int64, {})):    # near the 16-bit class limit
        x = synth_logits(shape, 77, dtype=dtype, device='cuda').requires_grad_(True)
        y = synth_labels((shape[0],) + shape[2:], C, 77, ignore_index=255, block=4, device='cuda').to(ldt)
        ce = B.CrossEntropyLoss(**kw)
        out = B.fused_resize_losses(x, y.unsqueeze(1), ce, ignore_index=255)
        out['loss_ce'].backward()
        xo = x.detach().float().requires_grad_(True)
        ref = O.cross_entropy_loss_module(xo, y.long(), ignore_index=255, **kw)
        ref.backward()
        acc = O.accuracy(xo.detach(), y.long(), ignore_index=255)
        lt, gt = (LOSS_TOL, GRAD_TOL) if dtype == torch.float32 else (HALF_TOL, 2 * HALF_TOL)
        assert rel_err(out['loss_ce'], ref) <= lt, '%s loss %.3e' % (shape, rel_err(out['loss_ce'], ref))
        assert rel_err(x.grad, xo.grad) <= gt, '%s grad %.3e' % (shape, rel_err(x.grad, xo.grad))
        assert abs(float(out['acc_seg']) - float(acc)) <= (1e-3 if dtype == torch.float32 else 0.5)



def test_cabi_output_buffers_have_no_out_of_bounds_writes(B):
    """Every output / workspace buffer of the single-pass entries sits between canary words (compute-sanitizer is not
    available on the GPU pool): resize-fused cell kernel + combine, and the bulk-copy pipeline with a partial last tile."""
    import ctypes as C
    from image_segmentation_lab_b200 import _lib
    lib = B.load_library()
    dev = torch.device('cuda', 0)
    stream = _lib.stream_ptr(dev)
    G = 4096                                    # guard floats on each side (16 KB: keeps the interior 16-byte aligned)

    def guarded(n_floats, dtype=torch.float32):
        buf = torch.full((n_floats + 2 * G,), 12345.0, dtype=torch.float32, device=dev)
        return buf, buf[G:G + n_floats]

    def check(buf, n_floats, what):
        assert bool((buf[:G] == 12345.0).all()) and bool((buf[G + n_floats:] == 12345.0).all()), what + ': guard overwritten'

    for (N, Cc, h, w, S) in ((2, 19, 9, 13, 8), (1, 5, 3, 2, 16), (3, 21, 6, 5, 4), (2, 21, 40, 52, 1), (1, 33, 20, 12, 1)):
        H, W = h * S, w * S
        x = synth_logits((N, Cc, h, w), 5, device='cuda')
        y = synth_labels((N, H, W), Cc, 5, ignore_index=255, block=3, device='cuda')
        gbuf, grad = guarded(N * Cc * h * w)
        sbuf, stats = guarded(16)
        nbytes = lib.b200seg_loss_fused_workspace_bytes(N, Cc, h, w, H, W, 0) if S > 1 else 0
        wbuf, ws = guarded(max(nbytes // 4, 4))
        fu = _lib.LossFusedDesc()
        fd = fu.fwd
        fd.logits = x.data_ptr(); fd.labels = y.data_ptr()
        fd.logit_dtype = _lib.F32; fd.label_dtype = _lib.L_I64
        fd.N, fd.C, fd.h, fd.w, fd.H, fd.W = N, Cc, h, w, H, W
        fd.flags = _lib.WANT_CE | _lib.WANT_ACC
        fd.ignore_index = 255; fd.acc_has_ignore = 1; fd.acc_ignore_index = 255
        fd.dice_exponent = 2.0; fd.ce_loss_weight = 1.0
        fd.stats = stats.data_ptr()
        fu.grad_scale_host = 1.0 / (N * H * W)
        fu.grad_logits = grad.data_ptr()
        fu.workspace = ws.data_ptr() if S > 1 else None
        _lib.check(lib.b200seg_loss_fused_fwdbwd(C.byref(fu), stream))
        torch.cuda.synchronize()
        check(gbuf, N * Cc * h * w, 'grad S=%d' % S)
        check(sbuf, 16, 'stats S=%d' % S)
        check(wbuf, max(nbytes // 4, 4), 'workspace S=%d' % S)
        # and the result is the oracle's gradient
        xo = x.clone().requires_grad_(True)
        full = O.resize(xo, size=(H, W), mode='bilinear', align_corners=False) if S > 1 else xo
        O.cross_entropy_loss_module(full, y, ignore_index=255).backward()
        assert rel_err(grad.view(N, Cc, h, w), xo.grad) <= GRAD_TOL


def test_config3_ade20k_shape(B):
    """BASELINE config 3 (batch reduced to 2 for the oracle's 150-iteration Python loop): 150 classes, 512x512, bf16,
    class-weighted CE + Dice(loss_weight=3)."""
    cw = torch.linspace(0.5, 1.5, 150).tolist()
    _head_case(B, (2, 150, 512, 512), (512, 512), 150, torch.bfloat16, dict(class_weight=cw), dict(loss_weight=3.0),
               loss_tol=HALF_TOL, grad_tol=2 * HALF_TOL)
    _head_case(B, (2, 150, 128, 128), (128, 128), 150, torch.float32, dict(class_weight=cw), dict(loss_weight=3.0))
    _head_case(B, (1, 150, 64, 64), (256, 256), 150, torch.float32, dict(class_weight=cw), dict(loss_weight=3.0))  # resize + dice


def test_config4_voc_shape(B):
    """BASELINE config 4 (batch 8 of the 32): 21 classes 512x512 fp32 CE, single-pass and two-pass plans."""
    _head_case(B, (8, 21, 512, 512), (512, 512), 21, torch.float32, {}, None)
    _head_case(B, (8, 21, 512, 512), (512, 512), 21, torch.float32, {}, None, single_pass=False)
    _head_case(B, (4, 21, 512, 512), (512, 512), 21, torch.float32, dict(avg_non_ignore=True), dict(class_weight=[1.0] * 21))


def test_ragged_shapes_and_unaligned_views(B):
    _head_case(B, (3, 7, 37, 53), (37, 53), 7, torch.float32, {}, dict())            # H*W odd -> scalar kernels
    _head_case(B, (1, 3, 1, 1), (1, 1), 3, torch.float32, {}, dict(), ignore=255, loss_atol=2e-7)    # a single pixel
    _head_case(B, (2, 33, 18, 22), (18, 22), 33, torch.float32, {}, dict())          # C just above one chunk
    _head_case(B, (1, 300, 16, 16), (16, 16), 300, torch.float32, {}, dict())        # 10 class groups x 30
    _head_case(B, (1, 600, 8, 8), (8, 8), 600, torch.float32, {}, None)              # CE only: any C
    # a non-contiguous / offset view must give the same answer as its contiguous copy
    big = synth_logits((2, 6, 20, 24), 3, device='cuda')
    view = big[:, 1:, 2:18, 1:17]
    labels = synth_labels((2, 16, 16), 5, 3, device='cuda', block=4)
    ce = B.CrossEntropyLoss()
    a = ce(view, labels, ignore_index=255)
    b = ce(view.contiguous(), labels, ignore_index=255)
    assert torch.equal(a, b)


def test_all_ignored_and_empty(B):
    x = synth_logits((2, 5, 16, 16), 1, device='cuda').requires_grad_(True)
    y = torch.full((2, 16, 16), 255, device='cuda')
    for kw in (dict(), dict(avg_non_ignore=True)):
        loss = B.CrossEntropyLoss(**kw)(x, y, ignore_index=255)
        ref = O.cross_entropy_loss_module(x.detach(), y, ignore_index=255, **kw)
        assert float(loss) == float(ref) == 0.0
        g, = torch.autograd.grad(loss, x)
        assert float(g.abs().max()) == 0.0
    acc = B.accuracy(x.detach(), y, ignore_index=255)
    assert abs(float(acc) - float(O.accuracy(x.detach(), y, ignore_index=255))) < 1e-4   # eps/eps * 100
    d = B.DiceLoss()(x, y)
    assert rel_err(d, O.dice_loss_module(x.detach(), y)) <= LOSS_TOL
    e = B.CrossEntropyLoss(reduction='sum')(torch.zeros(0, 5, 4, 4, device='cuda'), torch.zeros(0, 4, 4, dtype=torch.long, device='cuda'))
    assert float(e) == 0.0
    assert B.areas_device([], [], 5, 255).shape == (0, 3, 5)


def test_functional_and_2d_inputs(B):
    g = torch.Generator().manual_seed(4)
    p = torch.randn((33, 6), generator=g).cuda().requires_grad_(True)
    t = torch.randint(0, 6, (33,), generator=g).cuda()
    w = torch.rand((33,), generator=g).cuda()
    for kw in (dict(), dict(reduction='none'), dict(reduction='sum', class_weight=[1, 2, 3, 4, 5, 6.0])):
        kw_o = dict(kw)
        if 'class_weight' in kw_o:
            kw_o['class_weight'] = torch.tensor(kw_o['class_weight'], device='cuda')
        a = B.cross_entropy(p, t, weight=w, ignore_index=2, **kw)
        b = O.cross_entropy(p.detach(), t, weight=w, ignore_index=2, **kw_o)
        assert a.shape == b.shape and rel_err(a, b) <= LOSS_TOL
    x = synth_logits((2, 4, 8, 8), 2, device='cuda')
    y = synth_labels((2, 8, 8), 4, 2, device='cuda', block=2)
    assert rel_err(B.dice_loss(x, y, smooth=2, exponent=2), O.dice_loss_module(x, y, smooth=2)) <= LOSS_TOL


def test_amp_grad_scale_and_retain(B):
    """the upstream gradient is an arbitrary device scalar (GradScaler), never 1 by assumption."""
    x = synth_logits((2, 19, 16, 32), 5, device='cuda')
    y = synth_labels((2, 128, 256), 19, 5, device='cuda')
    for xx, size in ((x, (128, 256)), (synth_logits((2, 19, 128, 256), 6, device='cuda'), (128, 256))):
        xa = xx.clone().requires_grad_(True)
        r = B.fused_resize_losses(xa, y.unsqueeze(1), B.CrossEntropyLoss(), ignore_index=255)
        (r['loss_ce'] * 65536.0).backward()
        xb = xx.clone().requires_grad_(True)
        full = O.resize(xb, size=size, mode='bilinear', align_corners=False)
        (O.cross_entropy_loss_module(full, y, ignore_index=255) * 65536.0).backward()
        assert rel_err(xa.grad, xb.grad) <= GRAD_TOL
    # under fp16 autocast the loss comes back in float32, as ATen's autocast policy for cross_entropy does
    with torch.autocast('cuda', dtype=torch.float16):
        l = B.CrossEntropyLoss()(x.half(), synth_labels((2, 16, 32), 19, 5, device='cuda', block=4), ignore_index=255)
    assert l.dtype == torch.float32
    l = B.CrossEntropyLoss()(x.bfloat16(), synth_labels((2, 16, 32), 19, 5, device='cuda', block=4), ignore_index=255)
    assert l.dtype == torch.bfloat16


def test_decode_head_mixin(B):
    """losses() with the reference's signature and return layout (decode_head.py:261-321)."""
    import torch.nn as nn

    class Head(B.B200DecodeHeadLossMixin, nn.Module):
        def __init__(self):
            super().__init__()
            self.loss_decode = nn.ModuleList([B.CrossEntropyLoss(loss_weight=1.0), B.DiceLoss(loss_weight=3.0),
                                              B.CrossEntropyLoss(loss_weight=0.5, loss_name='loss_ce')])
            self.align_corners = False
            self.ignore_index = 255
            self.sampler = None

    head = Head()
    x = synth_logits((2, 6, 8, 8), 9, device='cuda').requires_grad_(True)
    y = synth_labels((2, 32, 32), 6, 9, device='cuda', block=4).unsqueeze(1)
    logits, loss = head.losses(x, y, {}, rescale=False)
    ref = O.head_losses(x.detach(), y, [('ce', dict(loss_weight=1.0), 'loss_ce'), ('dice', dict(loss_weight=3.0), 'loss_dice'),
                                        ('ce', dict(loss_weight=0.5), 'loss_ce')], ignore_index=255)
    assert list(loss.keys()) == ['loss_ce', 'loss_dice', 'acc_seg']
    for k in ref:
        assert rel_err(loss[k], ref[k]) <= (LOSS_TOL if k != 'acc_seg' else 1e-4), k
    assert logits.shape == x.shape
    infos = {'ori_img_size_hw': [(40, 36), (20, 28)]}
    resc, _ = head.losses(x.detach(), y, infos, rescale=True)
    full = O.resize(x.detach(), size=(32, 32), mode='bilinear', align_corners=False)
    for i, s in enumerate(infos['ori_img_size_hw']):
        want = O.resize(full[i].unsqueeze(0), size=s, mode='bilinear', align_corners=False)
        assert resc[i].shape == want.shape and rel_err(resc[i], want) <= 4e-6
    resc, _ = head.losses(x.detach(), y, {'ori_img_size_hw': (48, 48)}, rescale=True)
    assert rel_err(resc, O.resize(full, size=(48, 48), mode='bilinear', align_corners=False)) <= 4e-6


def test_cuda_graph_capture_replay(B):
    """Every entry point is stream-ordered and allocation-free inside the C ABI: the step captures and replays."""
    x0 = synth_logits((2, 19, 16, 32), 7, device='cuda')
    y = synth_labels((2, 128, 256), 19, 7, device='cuda').unsqueeze(1)
    ce = B.CrossEntropyLoss()

    def eager():  # own scope: no autograd node of the eager run may outlive it (AccumulateGrad is stream-bound)
        xe = x0.clone().requires_grad_(True)
        r = B.fused_resize_losses(xe, y, ce, ignore_index=255)
        r['loss_ce'].backward()
        return r['loss_ce'].detach().clone(), xe.grad.clone()

    want_loss, want_grad = eager()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        x = x0.clone().requires_grad_(True)
    torch.cuda.synchronize()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            r = B.fused_resize_losses(x, y, ce, ignore_index=255)
            r['loss_ce'].backward()
            x.grad = None
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        r = B.fused_resize_losses(x, y, ce, ignore_index=255)
        r['loss_ce'].backward()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(r['loss_ce'].detach(), want_loss) and torch.equal(x.grad, want_grad)


def test_determinism_bitwise(B):
    """The resize-fused backward is a fixed-order gather (ATen's is an atomicAdd scatter): runs are bit-identical."""
    x = synth_logits((4, 19, 32, 64), 8, device='cuda')
    y = synth_labels((4, 256, 512), 19, 8, device='cuda').unsqueeze(1)
    grads = []
    for _ in range(3):
        xa = x.clone().requires_grad_(True)
        B.fused_resize_losses(xa, y, B.CrossEntropyLoss(), ignore_index=255)['loss_ce'].backward()
        grads.append(xa.grad.clone())
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])


# ------------------------------------------------------------------------------------------------ evaluation at scale
def _areas_oracle_gpu(preds, gts, C, ignore):
    """Exact int64 areas with torch.bincount on the GPU (the numpy oracle's rule, for large inputs)."""
    out = []
    for p, g in zip(preds, gts):
        p = p.reshape(-1).long()
        g = g.reshape(-1).long()
        keep = g != ignore
        p, g = p[keep], g[keep]
        pin = (p >= 0) & (p < C)
        gin = (g >= 0) & (g < C)
        I = torch.bincount(p[(p == g) & pin], minlength=C)
        P = torch.bincount(p[pin], minlength=C)
        L = torch.bincount(g[gin], minlength=C)
        out.append(torch.stack([I, P, L]))
    return torch.stack(out)


@pytest.mark.parametrize('C', [2, 19, 60, 150, 400])
def test_areas_label_maps_all_counter_modes(B, C):
    """Private per-thread counters (C small), 128-thread variant (C mid) and shared atomics (C large)."""
    g = torch.Generator().manual_seed(C)
    sizes = [(257, 301), (64, 64), (1, 1), (511, 7), (1024, 2048)]
    preds = [torch.randint(-2, C + 3, s, generator=g).cuda() for s in sizes]
    gts = [synth_labels((1,) + s, C, C + i, ignore_index=255, block=16)[0].float().cuda() for i, s in enumerate(sizes)]
    gts[0][5, 5] = C + 7
    got = B.areas_device(preds, gts, C, 255)
    want = _areas_oracle_gpu(preds, gts, C, 255)
    assert torch.equal(got, want)
    small = O.intersect_and_union_int([p.cpu() for p in preds[:3]], [t.cpu() for t in gts[:3]], C, 255)
    assert np.array_equal(got[:3].cpu().numpy(), small[:, [0, 2, 3]])


def test_config5_sweep_properties(B):
    """BASELINE config 5 (64 of the 500 images on one GPU): 1024x2048, 19 classes; bit-exact against bincount, plus
    size-independent properties: sum(label) = #non-ignored, I <= min(P,L), U = P + L - I, additivity over shards."""
    C, n = 19, 64
    g = torch.Generator().manual_seed(55)
    gt_all = synth_labels((n, 1024, 2048), C, 55, ignore_index=255, block=16).float().cuda()
    pred_all = torch.randint(0, C, (n, 1024, 2048), generator=g, dtype=torch.int64).cuda()
    agree = torch.rand((n, 64, 128), generator=g).cuda().repeat_interleave(16, 1).repeat_interleave(16, 2) < 0.7
    pred_all = torch.where(agree & (gt_all != 255), gt_all.long(), pred_all)
    preds, gts = list(pred_all.unbind(0)), list(gt_all.unbind(0))
    a = B.areas_device(preds, gts, C, 255)
    assert torch.equal(a, _areas_oracle_gpu(preds, gts, C, 255))
    I, P, L = a[:, 0], a[:, 1], a[:, 2]
    valid = (gt_all != 255).sum(dim=(1, 2))
    assert torch.equal(L.sum(1), valid) and torch.equal(P.sum(1), valid)
    assert bool((I <= torch.minimum(P, L)).all())
    first = B.areas_device(preds[:40], gts[:40], C, 255).sum(0)
    second = B.areas_device(preds[40:], gts[40:], C, 255).sum(0)
    assert torch.equal(first + second, a.sum(0))
    # the same totals through the evaluator API + mIoU against the oracle's metric code on the exact totals
    lists = B.SegEvaluator.intersect_and_union(preds[:8], gts[:8], C, 255)
    tot = a[:8].sum(0)
    got = torch.stack([torch.stack(lists[j]).to(torch.int64).sum(0) for j in (0, 2, 3)]).cuda()
    assert torch.equal(got, tot)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_areas_from_logits(B, dtype):
    """process(): fused arg-max + areas vs softmax->argmax->histc of the reference on margin inputs."""
    C = 19
    shapes = [(1, C, 1024, 2048), (1, C, 300, 333), (1, C, 17, 5)]
    logits = [synth_logits(s, 60 + i, dtype=dtype, device='cuda') for i, s in enumerate(shapes)]
    gts = [synth_labels((1,) + s[2:], C, 60 + i, ignore_index=255)[0].float().cuda() for i, s in enumerate(shapes)]
    got = B.areas_device(logits, gts, C, 255, from_logits=True)
    preds = [O.argmax_labels(l.float()) for l in logits]
    assert torch.equal(got, _areas_oracle_gpu(preds, gts, C, 255))
    # un-margined inputs: report how often softmax rounding changes the arg-max (must stay tiny, never asserted to 0)
    raw = torch.randn((1, C, 512, 512), device='cuda').to(dtype)
    mism = int((O.argmax_labels(raw) != raw.float().argmax(1).squeeze(0)).sum())
    print('softmax->argmax vs argmax(logits) mismatches on un-margined %s logits: %d / %d' % (dtype, mism, 512 * 512))
    assert mism < 512 * 512 * 0.01


def test_ce_properties_full_size(B):
    """Size-independent properties at BASELINE config 2 / 4 sizes: shift invariance, zero gradient on ignored pixels,
    per-pixel gradient sums to zero over classes, and loss additivity over image shards."""
    x = synth_logits((8, 21, 512, 512), 70, device='cuda')
    y = synth_labels((8, 512, 512), 21, 70, device='cuda')
    ce = B.CrossEntropyLoss(reduction='sum')
    xa = x.clone().requires_grad_(True)
    l = ce(xa, y, ignore_index=255)
    l.backward()
    shift = torch.randn((8, 1, 512, 512), device='cuda')
    l2 = ce(x + shift, y, ignore_index=255)
    assert rel_err(l2, l) <= 1e-5
    g = xa.grad
    assert float(g.sum(1).abs().max()) <= 1e-5
    assert float(g[(y == 255).unsqueeze(1).expand_as(g)].abs().max()) == 0.0
    parts = sum(ce(x[i:i + 2], y[i:i + 2], ignore_index=255) for i in range(0, 8, 2))
    assert rel_err(parts, l) <= 1e-6


@pytest.mark.parametrize('ac', [False, True])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16])
def test_areas_from_lowres_logits_resize_fused(B, ac, dtype):
    """validation rescale + arg-max + areas in one kernel — low-resolution logits against full-resolution
    ground truth, never materialising the (1,C,H,W) rescaled logits (decode_head.py:297-320 + metrics.py:101-107).
    BIT-EXACT at every ratio and either align_corners setting: the kernel evaluates ATen's bilinear expression with its
    FMA contraction pinned (tools/probe/run_fma_probe.py), and rounds to the logit dtype as F.interpolate does."""
    C = 19
    cases = [((1, C, 64, 128), (512, 1024)), ((1, C, 32, 32), (256, 256)), ((1, C, 37, 53), (111, 160)), ((1, C, 9, 7), (9, 7)),
             ((1, C, 65, 129), (513, 1025)), ((1, C, 50, 70), (173, 301)), ((1, C, 7, 300), (40, 300))]
    g = torch.Generator().manual_seed(4242)
    # random (un-quantised) fp32 logits: no exact ties, every rounding of the interpolation matters
    logits = [(torch.randn(s, generator=g) * 3).to(dtype).cuda() for s, _ in cases]
    gts = [synth_labels((1,) + gt, C, 80 + i, ignore_index=255)[0].float().cuda() for i, (_, gt) in enumerate(cases)]
    maps = []
    got = B.areas_device(logits, gts, C, 255, from_logits=True, align_corners=ac, pred_maps=maps)
    full = [O.resize(l, size=tuple(gt.shape), mode='bilinear', align_corners=ac) for l, gt in zip(logits, gts)]
    assert all(f.dtype == dtype for f in full)
    preds = [f.float().argmax(dim=1).squeeze(0) for f in full]      # arg-max of the logits (lowest index wins ties)
    want = _areas_oracle_gpu(preds, gts, C, 255)
    for i, (p_, m_) in enumerate(zip(preds, maps)):
        assert torch.equal(p_, m_), (i, int((p_ != m_).sum()))
    assert torch.equal(got, want)
    tot = B.area_totals_device(logits, gts, C, 255, from_logits=True, align_corners=ac)
    assert torch.equal(tot, want.sum(0))
    if dtype == torch.float32:
        # margin inputs: the reference's softmax -> argmax (metrics.py:106) agrees with the arg-max of the logits
        lm = [synth_logits(s, 80 + i, device='cuda') for i, (s, _) in enumerate(cases[:4])]
        pm = [O.argmax_labels(O.resize(l, size=tuple(gt.shape), mode='bilinear', align_corners=ac)) for l, gt in zip(lm, gts[:4])]
        gm = B.areas_device(lm, gts[:4], C, 255, from_logits=True, align_corners=ac)
        assert torch.equal(gm, _areas_oracle_gpu(pm, gts[:4], C, 255))
        ev = B.SegEvaluator(epoch=0, num_classes=C, class_names=['c%d' % i for i in range(C)], palette=None, ignore_index=255,
                            show_result=False, align_corners=ac)
        ev.process(0, {'decode': [l.clone() for l in lm]}, {'ori_gt': gts[:4]})
        assert torch.equal(ev.area_totals('decode')[[0, 2, 3]], gm.sum(0).cpu())


@pytest.mark.parametrize('C', [19, 40, 300])
def test_areas_resize_fused_band_and_row_forms(B, C):
    """The two unit forms of confusion_resize_kernel (bands of up-sampled rows that share their horizontal sums; single
    rows otherwise) and both counter flavours (private columns for small C, shared atomics above; byte-packed class
    indices need C <= 255): long bands split into several groups, ragged last groups, rows down-sampled while columns
    are up-sampled, one-row and one-column logits, several images of different shapes in one launch — all bit-exact."""
    cases = [((1, C, 8, 16), (100, 129)), ((1, C, 5, 40), (64, 40)), ((1, C, 40, 6), (17, 90)), ((1, C, 1, 9), (13, 31)),
             ((1, C, 6, 1), (50, 7)), ((1, C, 33, 65), (264, 520)), ((1, C, 3, 3), (2, 2)), ((1, C, 16, 16), (17, 19))]
    g = torch.Generator().manual_seed(777 + C)
    logits = [(torch.randn(s, generator=g) * 3).cuda() for s, _ in cases]
    gts = [synth_labels((1,) + gt, min(C, 250), 300 + i, ignore_index=255)[0].float().cuda() for i, (_, gt) in enumerate(cases)]
    for ac in (False, True):
        maps = []
        got = B.areas_device(logits, gts, C, 255, from_logits=True, align_corners=ac, pred_maps=maps)
        preds = [O.resize(l, size=tuple(gt.shape), mode='bilinear', align_corners=ac).argmax(dim=1).squeeze(0)
                 for l, gt in zip(logits, gts)]
        for i, (p_, m_) in enumerate(zip(preds, maps)):
            assert torch.equal(p_, m_), (ac, i, int((p_ != m_).sum()))
        assert torch.equal(got, _areas_oracle_gpu(preds, gts, C, 255))
        assert torch.equal(B.area_totals_device(logits, gts, C, 255, from_logits=True, align_corners=ac), got.sum(0))
    # every ground-truth dtype the evaluator accepts goes through the same write-out (float32 above; int64 / uint8 have their
    # own vector forms, the others the scalar decoder), and 16-bit logits round the interpolated value before the compare
    want = _areas_oracle_gpu(preds, gts, C, 255)
    for gdt in (torch.int64, torch.uint8, torch.int32, torch.int16, torch.float64):
        got = B.areas_device(logits, [g_.to(gdt) for g_ in gts], C, 255, from_logits=True, align_corners=True)
        assert torch.equal(got, want), gdt
    for ldt in (torch.bfloat16, torch.float16):
        lh = [l.to(ldt) for l in logits]
        ph = [O.resize(l, size=tuple(gt.shape), mode='bilinear', align_corners=False).float().argmax(dim=1).squeeze(0)
              for l, gt in zip(lh, gts)]
        got = B.areas_device(lh, [g_.long() for g_ in gts], C, 255, from_logits=True, align_corners=False)
        assert torch.equal(got, _areas_oracle_gpu(ph, gts, C, 255)), ldt
