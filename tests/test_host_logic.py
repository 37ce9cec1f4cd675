"""CPU: host-side logic of the package — reduction helpers, registry shim, evaluator host maths, sharding and the
packed all-reduce (gloo, world_size 2), and the loud failure on non-CUDA inputs (there is no CPU fallback)."""
import os
import warnings
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import image_segmentation_lab_b200 as B
from image_segmentation_lab_b200 import distributed as D


def test_reduction_helpers_kat():
    """models/losses/utils.py:95-111 docstring example through this package's helpers."""
    l1 = B.weighted_loss(lambda pred, target: (pred - target).abs())
    pred, target, weight = torch.Tensor([0, 2, 3]), torch.Tensor([1, 1, 1]), torch.Tensor([1, 0, 1])
    assert abs(l1(pred, target).item() - 1.3333) < 1e-4
    assert l1(pred, target, weight).item() == 1.0
    assert l1(pred, target, reduction='none').tolist() == [1.0, 1.0, 2.0]
    assert abs(l1(pred, target, weight, avg_factor=2).item() - 1.5) < 1e-6
    with pytest.raises(ValueError):
        l1(pred, target, weight, reduction='sum', avg_factor=2)
    assert B.get_class_weight([1, 2]) == [1, 2]


def test_modules_are_stateless_and_named():
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        ce = B.CrossEntropyLoss(class_weight=[1.0, 2.0], loss_weight=0.4, loss_name='loss_aux')
        assert any('avg_non_ignore' in str(x.message) for x in w)   # cross_entropy_loss.py:244-249
    dice = B.DiceLoss(loss_weight=3.0)
    assert ce.loss_name == 'loss_aux' and dice.loss_name == 'loss_dice'
    assert len(ce.state_dict()) == 0 and len(dice.state_dict()) == 0      # checkpoint compatible
    assert 'avg_non_ignore=False' in repr(ce)
    with pytest.raises(AssertionError):
        B.CrossEntropyLoss(use_sigmoid=True, use_mask=True)


def test_no_cpu_fallback():
    ce = B.CrossEntropyLoss(avg_non_ignore=True)
    x, y = torch.randn(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long)
    with pytest.raises(RuntimeError, match='no CPU'):
        ce(x, y)
    with pytest.raises(RuntimeError, match='no CPU'):
        B.resize(x, size=(8, 8), mode='bilinear')
    with pytest.raises(RuntimeError, match='no CPU'):
        B.accuracy(x, y)


def test_argument_validation_mirrors_reference():
    ce = B.CrossEntropyLoss(avg_non_ignore=True)
    x, y = torch.randn(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long)
    with pytest.raises(AssertionError):                      # cross_entropy_loss.py:273
        ce(x, y, reduction_override='avg')
    with pytest.raises(ValueError):                          # utils.py:78-79
        ce(x, y, avg_factor=3.0, reduction_override='sum')
    with pytest.raises(AssertionError):                      # accuracy.py:38
        B.accuracy(x, y, topk=5)
    with pytest.raises(KeyError):                            # metrics.py:319-320
        B.SegEvaluator.total_area_to_metrics(torch.ones(2), torch.ones(2), torch.ones(2), torch.ones(2), ['mAP'])


def test_registry_install_and_build():
    class FakeLoss:                                          # shape of registry/register.py:9-28
        _storage = {'CrossEntropyLoss': object, 'DiceLoss': object}

        @classmethod
        def get(cls, name):
            if name not in cls._storage:
                raise KeyError(name)
            return cls._storage[name]

    installed = B.registry.install(FakeLoss, override=True)
    assert FakeLoss.get('CrossEntropyLoss') is B.CrossEntropyLoss and FakeLoss.get('B200DiceLoss') is B.DiceLoss
    assert set(installed) == {'CrossEntropyLoss', 'DiceLoss', 'TverskyLoss', 'LovaszLoss', 'B200CrossEntropyLoss', 'B200DiceLoss',
                              'B200TverskyLoss', 'B200LovaszLoss'}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = B.registry.build_loss(dict(type='CrossEntropyLoss', loss_weight=0.4, class_weight=[1.0, 2.0]))
    assert isinstance(m, B.CrossEntropyLoss) and m.loss_weight == 0.4
    with pytest.raises(KeyError):
        B.registry.build_loss(dict(type='FocalLoss'))
    with pytest.raises(AssertionError):
        B.registry.build_loss(dict(type='LovaszLoss'))   # per_image=False needs reduction='none' (lovasz_loss.py:271-273)
    with pytest.raises(TypeError):
        B.registry.build_loss('CrossEntropyLoss')


def test_total_area_to_metrics_matches_reference_fixture(golden):
    data, _ = golden
    I, U, P, L = (torch.from_numpy(data['metrics/' + k]) for k in 'IUPL')
    for tag, kw in [('plain', {}), ('nan0_beta2', dict(nan_to_num=0, beta=2))]:
        r = B.SegEvaluator.total_area_to_metrics(I, U, P, L, ['mIoU', 'mDice', 'mFscore'], **kw)
        assert list(r.keys()) == ['aAcc', 'IoU', 'Acc', 'Dice', 'Fscore', 'Precision', 'Recall']
        for k, v in r.items():
            np.testing.assert_array_equal(v, data['metrics_%s/%s' % (tag, k)])


def test_seg_metrics_from_prefilled_results(golden, capsys):
    """compute_metrics on the per-image areas the reference produced (host maths only)."""
    data, _ = golden
    areas = torch.from_numpy(data['process/areas'])          # (n,4,C) float32
    ev = B.SegEvaluator(epoch=0, num_classes=5, class_names=['c%d' % i for i in range(5)], palette=None,
                        ignore_index=-1, show_result=False)
    ev.results = {'decode': [list(areas[:, j].unbind(0)) for j in range(4)]}
    met = ev.compute_metrics()['decode']
    for k in ('aAcc', 'mIoU', 'mAcc', 'mDice', 'mFscore', 'mPrecision', 'mRecall'):
        assert met[k] == data['process/summary_' + k], k
    for k in ('IoU', 'Acc', 'Dice', 'Fscore', 'Precision', 'Recall'):
        np.testing.assert_array_equal(met[k], data['process/class_' + k])
    assert met['Class'] == ['c%d' % i for i in range(5)]
    assert 'decode' in capsys.readouterr().out
    ev2 = B.SegEvaluator(epoch=0, num_classes=5, class_names=['c%d' % i for i in range(5)], palette=None,
                         ignore_index=-1, show_result=False, exact_totals=False)
    ev2.results = {'decode': [list(areas[:, j].unbind(0)) for j in range(4)]}
    assert ev2.compute_metrics()['decode']['mIoU'] == met['mIoU']


def test_exact_totals_beat_fp32_accumulation():
    """500 x (2097152 + 1) summed in fp32 loses the +1s; the int64 totals do not."""
    n = 500
    per = torch.full((1,), 2097152.0)
    lst = [per.clone() for _ in range(n)]
    lst[0] = per + 1
    assert int(sum(lst).item()) == 1048576000                  # the reference's fp32 running sum
    exact = torch.stack([t.to(torch.int64) for t in lst]).sum(0)
    assert int(exact.item()) == 1048576001


def test_shard_range_covers_everything():
    for n, w in [(500, 8), (32, 8), (7, 4), (3, 8), (0, 2)]:
        seen = []
        for r in range(w):
            lo, hi = D.shard_range(n, r, w)
            assert 0 <= lo <= hi <= n
            seen += list(range(lo, hi))
        assert seen == list(range(n))
    assert D.shard_range(500, 0, 8) == (0, 63) and D.shard_range(500, 7, 8) == (441, 500)


def test_global_loss_scalars():
    vec = torch.tensor([120.0, 90.0, 45.0, 90.0, 0.0, 100.0, 0.0, 2.0], dtype=torch.float64)
    loss, acc = D.global_loss_scalars(vec, loss_weight=0.5)
    assert abs(loss.item() - 0.6) < 1e-6 and abs(acc.item() - 50.0) < 1e-4
    loss2, _ = D.global_loss_scalars(vec, avg_non_ignore=True)
    assert abs(loss2.item() - 120.0 / 90.0) < 1e-6


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    D.init_from_env(backend='gloo')
    try:
        # each rank owns a contiguous image range; areas are exact integers > 2**24
        n_img, C = 7, 5
        lo, hi = D.shard_range(n_img, rank, world)
        g = torch.Generator().manual_seed(1)
        all_areas = torch.randint(0, 2 ** 40, (n_img, 4, C), generator=g, dtype=torch.int64)
        local = {'decode': all_areas[lo:hi].sum(0), 'aux': all_areas[lo:hi].sum(0) * 2}
        red = D.all_reduce_areas(local)
        ok = torch.equal(red['decode'], all_areas.sum(0)) and torch.equal(red['aux'], all_areas.sum(0) * 2)
        ok = ok and red['decode'].dtype == torch.int64
        # packed loss scalars: [ce_sum, n_valid, n_correct, n_acc, n_pixels]
        par = D.PackedAllReduce()
        vec = torch.tensor([10.0 * (rank + 1), 100, 50 + rank, 100, 0, 128, 0, 2], dtype=torch.float64)
        cnt = torch.tensor([3 + rank], dtype=torch.int64)
        par.start([vec, cnt])
        rv, rc = par.finish()
        ok = ok and abs(rv[0].item() - 30.0) < 1e-12 and rv[5].item() == 256 and rc.item() == 7 and rc.dtype == torch.int64
        loss, acc = D.global_loss_scalars(rv)
        ok = ok and abs(loss.item() - 30.0 / 256) < 1e-7
        # parse_losses: one all-reduce for every logged variable, rank-mean as the reference (train_utils.py:69-72)
        _, lv = B.parse_losses({'loss_ce': torch.tensor(1.0 + rank), 'acc_seg': torch.tensor([10.0 * rank])})
        ok = ok and abs(lv['loss_ce'] - 1.5) < 1e-6 and abs(lv['acc_seg'] - 5.0) < 1e-6 and abs(lv['loss'] - 1.5) < 1e-6
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_packed_allreduce_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_parse_losses_matches_reference_layout():
    """utils/train_utils.py:31-74: keys containing 'loss' are summed into the graph; everything is logged as floats."""
    a = torch.tensor(1.5, requires_grad=True)
    b = torch.tensor([0.25, 0.75], requires_grad=True)
    losses = OrderedDict([('decode.loss_ce', a * 2), ('decode.acc_seg', torch.tensor([87.5])),
                          ('aux.loss_ce', [b[0] * 1.0, b[1] * 1.0]), ('_stats', torch.zeros(8))])
    loss, log_vars = B.parse_losses(losses)
    assert list(log_vars.keys()) == ['decode.loss_ce', 'decode.acc_seg', 'aux.loss_ce', 'loss']
    assert abs(log_vars['loss'] - 4.0) < 1e-6 and abs(log_vars['decode.acc_seg'] - 87.5) < 1e-6
    assert all(isinstance(v, float) for v in log_vars.values())
    loss.backward()
    assert float(a.grad) == 2.0 and b.grad.tolist() == [1.0, 1.0]
    _, lazy = B.parse_losses(OrderedDict([('loss_x', torch.tensor(2.0))]), lazy=True)
    assert isinstance(lazy['loss'], torch.Tensor) and float(lazy['loss']) == 2.0
    with pytest.raises(TypeError):
        B.parse_losses({'loss_bad': 1.0})


def test_packed_allreduce_single_process_is_identity():
    par = D.PackedAllReduce()
    a, b = torch.arange(3, dtype=torch.int64), torch.tensor([1.5])
    out = par([a, b])
    assert torch.equal(out[0], a) and torch.equal(out[1], b)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the CUDA arm): one JSON line with the contract's keys,
    the headline metric / workload of the CUDA arm, and the reference's own files as the thing timed wherever they can be
    executed (oracle/_ref byte code or /root/reference), the oracle port otherwise."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in line, k
    assert line['impl'] == 'reference' and line['unit'] == 'Mpix/s' and line['value'] > 0
    assert line['config']['batch_per_gpu'] == 8 and 'workload' in line['config']
    cb = line['cpu_baseline']
    from oracle import ref_loader
    assert cb['kind'] == ('reference' if ref_loader.available() else 'port')
    assert cb['value'] == line['value'] and cb['cores'] >= 1 and '8 of the 8 images' in cb['sample']
    assert line['e2e'] == {'value': line['value'], 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    # under torchrun only rank 0 prints; the other ranks exit 0 without work
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    r1 = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                         '--warmup', '0'], capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert r1.returncode == 0 and r1.stdout.strip() == ''
