"""CPU: the oracle restatement reproduces the fixtures recorded from the reference's own files."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.helpers import case_tensors, loss_case_oracle

RT = dict(rtol=2e-6, atol=1e-7)  # same ATen ops; tolerance only for CPU vector-width differences across hosts


def _loss_cases(manifest):
    return [c for c in manifest['cases'] if c['kind'] in ('ce', 'dice', 'head')]


def test_manifest_has_cases(golden):
    data, manifest = golden
    assert len(_loss_cases(manifest)) >= 20
    assert manifest['torch'].startswith('2.')


def test_loss_cases(golden):
    data, manifest = golden
    for case in _loss_cases(manifest):
        name = case['name']
        logits, labels, pw = case_tensors(data, name)
        go = torch.from_numpy(data[name + '/grad_out']) if (name + '/grad_out') in data else None
        out = loss_case_oracle(case, logits, labels, pw, go)
        for k, v in out.items():
            np.testing.assert_allclose(v.numpy(), data[name + '/' + k], err_msg='%s/%s' % (name, k), **RT)


def test_kat_docstring_example():
    """models/losses/utils.py:95-111 — the reference's only known-answer values."""
    pred, target, weight = torch.Tensor([0, 2, 3]), torch.Tensor([1, 1, 1]), torch.Tensor([1, 0, 1])
    loss = (pred - target).abs()
    assert abs(O.weight_reduce_loss(loss).item() - 1.3333) < 1e-4
    assert O.weight_reduce_loss(loss, weight).item() == 1.0
    assert O.weight_reduce_loss(loss, reduction='none').tolist() == [1.0, 1.0, 2.0]
    assert abs(O.weight_reduce_loss(loss, weight, avg_factor=2).item() - 1.5) < 1e-6
    with pytest.raises(ValueError):
        O.weight_reduce_loss(loss, weight, reduction='sum', avg_factor=2)


def test_kat_matches_fixture(golden):
    data, _ = golden
    np.testing.assert_allclose(data['kat/out'], [4.0 / 3.0, 1.0, 1.5], rtol=1e-6)
    np.testing.assert_array_equal(data['kat/none'], [1.0, 1.0, 2.0])


def test_resize(golden):
    data, manifest = golden
    for case in [c for c in manifest['cases'] if c['kind'] == 'resize']:
        name = case['name']
        x = torch.from_numpy(data[name + '/x']).requires_grad_(True)
        y = O.resize(x, size=tuple(case['size']), mode='bilinear', align_corners=case['ac'])
        y.backward(torch.from_numpy(data[name + '/go']))
        np.testing.assert_allclose(y.detach().numpy(), data[name + '/y'], **RT)
        np.testing.assert_allclose(x.grad.numpy(), data[name + '/gx'], rtol=1e-5, atol=1e-6)
    y = O.resize(torch.from_numpy(data['resize_nearest/x']), size=(9, 15))
    np.testing.assert_array_equal(y.numpy(), data['resize_nearest/y'])


def test_accuracy_topk(golden):
    data, _ = golden
    r = O.accuracy(torch.from_numpy(data['acc_topk/pred']), torch.from_numpy(data['acc_topk/target']), topk=(1, 3), thresh=0.2)
    np.testing.assert_allclose(np.stack([v.numpy() for v in r]), data['acc_topk/out'], rtol=1e-6)
    r = O.accuracy(torch.from_numpy(data['acc_topk4d/pred']), torch.from_numpy(data['acc_topk4d/target']), topk=(1, 2, 5),
                   ignore_index=255)
    np.testing.assert_allclose(np.stack([v.numpy() for v in r]), data['acc_topk4d/out'], rtol=1e-6)


def test_intersect_and_union_bit_exact(golden):
    data, _ = golden
    preds = [torch.from_numpy(data['iau/pred%d' % i]) for i in range(3)]
    gts = [torch.from_numpy(data['iau/gt%d' % i]) for i in range(3)]
    a = O.intersect_and_union_int(preds, gts, 5, 255)
    np.testing.assert_array_equal(a, data['iau/areas'].astype(np.int64))
    lists = O.intersect_and_union(preds, gts, 5, 255)
    assert len(lists) == 4 and lists[0][0].dtype == torch.float32
    t = O.intersect_and_union_torch(preds, gts, 5, 255)
    for j in range(4):
        np.testing.assert_array_equal(np.stack([x.numpy() for x in t[j]]), data['iau/areas'][:, j])


def test_process_argmax_and_metrics(golden):
    data, _ = golden
    logits = [torch.from_numpy(data['process/logits%d' % i]) for i in range(3)]
    gts = [torch.from_numpy(data['process/gt%d' % i]) for i in range(3)]
    preds = [O.argmax_labels(l) for l in logits]
    a = O.intersect_and_union_int(preds, gts, 5, -1)
    np.testing.assert_array_equal(a, data['process/areas'].astype(np.int64))
    tot = torch.from_numpy(a.sum(0)).to(torch.float32)
    ret = O.total_area_to_metrics(tot[0], tot[1], tot[2], tot[3], ['mIoU', 'mDice', 'mFscore'])
    summ = O.summarize(ret)
    for k in ('aAcc', 'mIoU', 'mAcc', 'mDice', 'mFscore', 'mPrecision', 'mRecall'):
        assert summ[k] == data['process/summary_' + k], k
    for k in ('IoU', 'Acc', 'Dice', 'Fscore', 'Precision', 'Recall'):
        np.testing.assert_array_equal(np.round(ret[k] * 100, 2), data['process/class_' + k])


def test_total_area_to_metrics_nan_and_beta(golden):
    data, _ = golden
    I, U, P, L = (torch.from_numpy(data['metrics/' + k]) for k in 'IUPL')
    for tag, kw in [('plain', {}), ('nan0_beta2', dict(nan_to_num=0, beta=2))]:
        r = O.total_area_to_metrics(I, U, P, L, ['mIoU', 'mDice', 'mFscore'], **kw)
        for k, v in r.items():
            np.testing.assert_array_equal(v, data['metrics_%s/%s' % (tag, k)])
    with pytest.raises(KeyError):
        O.total_area_to_metrics(I, U, P, L, ['mAP'])


def test_bce_cases(golden):
    """Sigmoid path: oracle.binary_cross_entropy vs CrossEntropyLoss(use_sigmoid=True) of the reference."""
    data, manifest = golden
    cases = [c for c in manifest['cases'] if c['kind'] == 'bce']
    assert len(cases) >= 6
    for case in cases:
        name, kw = case['name'], dict(case['kw'])
        x = torch.from_numpy(data[name + '/logits']).requires_grad_(True)
        y = torch.from_numpy(data[name + '/labels'])
        w = torch.from_numpy(data[name + '/pixel_weight']) if case['pixel_weight'] else None
        if 'class_weight' in kw:
            kw['class_weight'] = x.new_tensor(kw['class_weight'])
        loss = case['loss_weight'] * O.binary_cross_entropy(x, y, w, ignore_index=255, **kw)
        if loss.dim():
            (loss * torch.from_numpy(data[name + '/grad_out'])).sum().backward()
        else:
            loss.backward()
        np.testing.assert_allclose(loss.detach().numpy(), data[name + '/loss'], err_msg=name, **RT)
        np.testing.assert_allclose(x.grad.numpy(), data[name + '/grad'], err_msg=name, rtol=1e-5, atol=1e-8)


def test_tversky_cases(golden):
    """TverskyLoss (models/losses/tversky_loss.py:24-148): oracle.tversky_loss_module vs the reference's fixtures."""
    data, manifest = golden
    cases = [c for c in manifest['cases'] if c['kind'] == 'tversky']
    assert len(cases) >= 4
    for case in cases:
        name = case['name']
        x = torch.from_numpy(data[name + '/logits']).requires_grad_(True)
        y = torch.from_numpy(data[name + '/labels'])
        loss = O.tversky_loss_module(x, y, **case['kw'])
        loss.backward()
        np.testing.assert_allclose(loss.detach().numpy(), data[name + '/loss'], err_msg=name, **RT)
        np.testing.assert_allclose(x.grad.numpy(), data[name + '/grad'], err_msg=name, rtol=1e-5, atol=1e-8)


def _lovasz_case(data, case, acc_dtype, dtype=torch.float32):
    name = case['name']
    x = torch.from_numpy(data[name + '/logits']).to(dtype).requires_grad_(True)
    y = torch.from_numpy(data[name + '/labels'])
    loss = O.lovasz_loss_module(x, y, avg_factor=case.get('avg_factor'), ignore_index=case['ignore'], acc_dtype=acc_dtype,
                                **case['kw'])
    if loss.dim():
        (loss * torch.from_numpy(data[name + '/grad_out']).to(dtype)).sum().backward()
    else:
        loss.backward()
    return loss.detach(), x.grad


def test_lovasz_cases(golden):
    """LovaszLoss (models/losses/lovasz_loss.py:26-312): the fp32 restatement reproduces the reference's fixtures; the
    exact (float64 Jaccard) restatement — what the CUDA path is held to — agrees with them within the gates."""
    data, manifest = golden
    cases = [c for c in manifest['cases'] if c['kind'] == 'lovasz']
    assert len(cases) >= 12
    for case in cases:
        name = case['name']
        assert case['order_margin'] > 1.0
        loss, grad = _lovasz_case(data, case, torch.float32)
        np.testing.assert_allclose(loss.numpy(), data[name + '/loss'], err_msg=name, **RT)
        np.testing.assert_allclose(grad.numpy(), data[name + '/grad'], err_msg=name, rtol=1e-5, atol=1e-8)
        loss64, grad64 = _lovasz_case(data, case, torch.float64, torch.float64)
        assert np.abs(loss64.numpy() - data[name + '/loss']).max() <= 1e-5 * np.abs(data[name + '/loss']).max(), name
        assert np.abs(grad64.numpy() - data[name + '/grad']).max() <= 1e-4 * np.abs(data[name + '/grad']).max(), name


def test_lovasz_grad_closed_form():
    """The closed form the CUDA kernel uses for the Jaccard increments (csrc/loss_lovasz.cu) equals lovasz_grad (:26-39)."""
    g = torch.Generator().manual_seed(11)
    for n, p_fg in ((1, 1.0), (1, 0.0), (7, 0.5), (300, 0.1), (300, 0.0), (2000, 0.7)):
        fg = (torch.rand(n, generator=g) < p_fg).double()
        ref = O.lovasz_grad(fg, torch.float64)
        gts = fg.sum()
        cum = fg.cumsum(0)
        i = torch.arange(n, dtype=torch.float64)
        I, U = gts - cum, gts + (i + 1 - cum)
        mine = torch.where(fg > 0, 1.0 / U, I / (U * (U - 1)).clamp(min=1))
        mine[0] = 1.0 - I[0] / U[0]
        np.testing.assert_allclose(mine.numpy(), ref.numpy(), rtol=1e-9, atol=1e-12)
