"""CPU, build container only: the oracle against the LIVE reference files on fresh random inputs."""
import warnings

import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='neither /root/reference nor oracle/_ref byte code on this machine')


@pytest.fixture(scope='module')
def ref():
    warnings.simplefilter('ignore')
    return ref_loader.load()


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_head_chain_bitwise(ref, seed):
    g = torch.Generator().manual_seed(seed)
    n, c, h, w, H, W = 2, 6, 5, 7, 20, 28
    x = torch.randn((n, c, h, w), generator=g)
    y = torch.randint(0, c, (n, 1, H, W), generator=g)
    y[torch.rand((n, 1, H, W), generator=g) < 0.2] = 255
    cw = (torch.rand(c, generator=g) + 0.5).tolist()
    for ac in (False, True):
        xr = x.clone().requires_grad_(True)
        full = ref.resize(xr, size=(H, W), mode='bilinear', align_corners=ac, warning=False)
        l1 = ref.CrossEntropyLoss(class_weight=cw, avg_non_ignore=bool(seed % 2))(full, y.squeeze(1), ignore_index=255)
        l2 = ref.DiceLoss(loss_weight=3.0, class_weight=cw)(full, y.squeeze(1), ignore_index=255)
        acc = ref.accuracy(full, y.squeeze(1), ignore_index=255)
        (l1 + l2).backward()
        xo = x.clone().requires_grad_(True)
        o = O.head_losses(xo, y, [('ce', dict(class_weight=cw, avg_non_ignore=bool(seed % 2)), 'loss_ce'),
                                  ('dice', dict(loss_weight=3.0, class_weight=cw), 'loss_dice')], align_corners=ac,
                          ignore_index=255)
        (o['loss_ce'] + o['loss_dice']).backward()
        assert torch.equal(o['loss_ce'], l1) and torch.equal(o['loss_dice'], l2) and torch.equal(o['acc_seg'], acc)
        assert torch.equal(xo.grad, xr.grad)


def test_binary_cross_entropy_bitwise(ref):
    g = torch.Generator().manual_seed(5)
    x = torch.randn((2, 3, 6, 6), generator=g)
    y = torch.randint(0, 3, (2, 6, 6), generator=g)
    y[0, 0, :2] = 255
    for kw in (dict(), dict(reduction='sum'), dict(avg_non_ignore=True)):
        a = ref.binary_cross_entropy(x, y, ignore_index=255, **kw)
        b = O.binary_cross_entropy(x, y, ignore_index=255, **kw)
        assert torch.equal(a, b)


def test_intersect_and_union_and_metrics(ref):
    g = torch.Generator().manual_seed(3)
    C = 19
    preds = [torch.randint(0, C, (33, 47), generator=g) for _ in range(4)]
    gts = [torch.randint(0, C, (33, 47), generator=g).float() for _ in range(4)]
    for t in gts:
        t[torch.rand(t.shape, generator=g) < 0.1] = 255
    r = ref_loader.intersect_and_union_cpu(ref, preds, gts, C, 255)
    a = O.intersect_and_union_int(preds, gts, C, 255)
    for j in range(4):
        np.testing.assert_array_equal(np.stack([x.numpy() for x in r[j]]).astype(np.int64), a[:, j])
    tot = [torch.from_numpy(a[:, j].sum(0)).float() for j in range(4)]
    rm = ref.SegEvaluator.total_area_to_metrics(*tot, ['mIoU', 'mDice', 'mFscore'], None, 1)
    om = O.total_area_to_metrics(*tot, ['mIoU', 'mDice', 'mFscore'])
    for k in rm:
        np.testing.assert_array_equal(rm[k], om[k])


def test_argmax_matches_process(ref):
    g = torch.Generator().manual_seed(11)
    l = torch.randn((1, 7, 12, 9), generator=g)
    assert torch.equal(O.argmax_labels(l), torch.nn.functional.softmax(l, dim=1).argmax(dim=1).squeeze(0))


@pytest.mark.parametrize('kw', [dict(reduction='none'), dict(reduction='none', classes='all', class_weight=[.5, 1, 1.5, 2, .25]),
                                dict(reduction='none', classes=[0, 2]), dict(per_image=True), dict(per_image=True, reduction='sum'),
                                dict(loss_type='binary', reduction='none'), dict(loss_type='binary', per_image=True)])
def test_lovasz_bitwise(ref, kw):
    g = torch.Generator().manual_seed(9)
    binary = kw.get('loss_type') == 'binary'
    x = (torch.randn((3, 1 if binary else 5, 7, 9), generator=g) * 2)
    y = torch.randint(0, 2 if binary else 5, (3, 7, 9), generator=g)
    y[torch.rand(y.shape, generator=g) < 0.15] = 255
    xr = x.clone().requires_grad_(True)
    a = ref.LovaszLoss(**kw)(xr, y, ignore_index=255)
    a.sum().backward()
    xo = x.clone().requires_grad_(True)
    b = O.lovasz_loss_module(xo, y, ignore_index=255, **kw)
    b.sum().backward()
    assert torch.equal(a, b) and torch.equal(xr.grad, xo.grad)


@pytest.mark.skipif(not ref_loader.source_available(), reason='needs the reference sources to compile')
def test_bytecode_form_is_the_same_code(tmp_path, monkeypatch):
    """oracle/_ref (what travels to the GPU box) is the reference's code: compiled from the sources where they lie, every
    function's byte code equals the source module's, and a fresh load from byte code alone runs the head chain."""
    from oracle import build_ref
    assert build_ref.build() == len(build_ref.FILES) and build_ref.usable()
    src = ref_loader.load()
    monkeypatch.setattr(ref_loader, 'REF_ROOT', str(tmp_path))
    monkeypatch.setattr(ref_loader, '_CACHE', None)
    assert ref_loader.origin() == 'bytecode'
    bc = ref_loader.load()
    for name in ('resize', 'cross_entropy', 'binary_cross_entropy', 'accuracy', 'weight_reduce_loss', 'lovasz_grad'):
        assert getattr(bc, name).__code__.co_code == getattr(src, name).__code__.co_code, name
    for name in ('CrossEntropyLoss', 'DiceLoss', 'TverskyLoss', 'LovaszLoss'):
        assert getattr(bc, name).forward.__code__.co_code == getattr(src, name).forward.__code__.co_code, name
    g = torch.Generator().manual_seed(11)
    x = torch.randn((2, 5, 6, 7), generator=g)
    y = torch.randint(0, 5, (2, 12, 14), generator=g)
    a = src.CrossEntropyLoss()(src.resize(x, size=(12, 14), mode='bilinear', align_corners=False), y, ignore_index=255)
    b = bc.CrossEntropyLoss()(bc.resize(x, size=(12, 14), mode='bilinear', align_corners=False), y, ignore_index=255)
    assert torch.equal(a, b)
