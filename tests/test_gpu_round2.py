"""GPU: round-2 parity cases — AMP (float16 + GradScaler) gradients on label-resolution logits, the stock head's in-place
`loss[name] += ...` flow, several devices in one process, and the multi-GPU (NCCL) sharded totals.

Gates as in test_gpu_parity.py."""
import os
import subprocess
import sys
import warnings

import pytest
import torch

from oracle import oracle as O
from tests.helpers import rel_err, synth_labels, synth_logits

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOSS_TOL = 1e-5
GRAD_TOL = 1e-4
HALF_TOL = 2.0 ** -7


@pytest.fixture(scope='module')
def B():
    import image_segmentation_lab_b200 as pkg
    pkg.load_library()
    warnings.simplefilter('ignore')
    return pkg


def test_fp16_gradscaler_gradient_keeps_its_bits(B):
    """The reference's shipped schedule is fp16 autocast + GradScaler (configs/schedule/kvasir_training_schedule.py:22,
    utils/train_utils.py:85-91): the upstream gradient is 65536 and autograd applies it in fp32 before the single cast
    to float16. A gradient formed at loss_weight/(N*H*W) ~ 5e-7 and scaled afterwards would be subnormal float16."""
    for shape, C in (((2, 19, 128, 256), 19), ((2, 21, 64, 96), 21), ((1, 60, 64, 64), 60)):
        x = synth_logits(shape, 11, device='cuda').half()
        y = synth_labels((shape[0],) + shape[2:], C, 11, device='cuda', block=8)
        xa = x.clone().requires_grad_(True)
        with torch.autocast('cuda', dtype=torch.float16):   # the loss comes back in float32, as ATen's autocast policy does
            la = B.CrossEntropyLoss()(xa, y, ignore_index=255)
        assert la.dtype == torch.float32
        (la * 65536.0).backward()
        xb = x.float().requires_grad_(True)
        lb = O.cross_entropy_loss_module(xb, y, ignore_index=255)
        (lb * 65536.0).backward()
        assert xa.grad.dtype == torch.float16
        # float16 has 11 significand bits: 2**-10 relative to the largest entry, and small entries must not collapse
        assert rel_err(xa.grad, xb.grad) <= 2.0 ** -10, rel_err(xa.grad, xb.grad)
        ref = xb.grad
        big = ref.abs() > 1e-3 * ref.abs().max()
        ratio = (xa.grad.float()[big] / ref[big])
        assert float((ratio - 1).abs().max()) <= 2.0 ** -9, float((ratio - 1).abs().max())


def test_same_named_losses_accumulate_in_place(B):
    """decode_head.py:283-293: `loss[name] = ...` then `loss[name] += ...` for losses sharing a loss_name. The loss
    scalars must therefore be tensors of their own (autograd refuses in-place updates of views made by a Function)."""
    x = synth_logits((2, 6, 32, 32), 3, device='cuda').requires_grad_(True)
    y = synth_labels((2, 32, 32), 6, 3, device='cuda', block=4)
    mods = [B.CrossEntropyLoss(loss_weight=1.0), B.DiceLoss(loss_weight=3.0, loss_name='loss_ce'),
            B.CrossEntropyLoss(loss_weight=0.5, loss_name='loss_ce'), B.LovaszLoss(reduction='none', loss_name='loss_ce'),
            B.TverskyLoss(loss_name='loss_ce')]
    loss = dict()
    for m in mods:                                   # the reference's loop, verbatim semantics
        if m.loss_name not in loss:
            loss[m.loss_name] = m(x, y, ignore_index=255)
        else:
            loss[m.loss_name] += m(x, y, ignore_index=255)
    loss['loss_ce'].backward()
    xo = x.detach().clone().requires_grad_(True)
    ref = (O.cross_entropy_loss_module(xo, y, ignore_index=255) + O.dice_loss_module(xo, y, loss_weight=3.0)
           + O.cross_entropy_loss_module(xo, y, ignore_index=255, loss_weight=0.5)
           + O.lovasz_loss_module(xo, y, reduction='none', ignore_index=255)
           + O.tversky_loss_module(xo, y, ignore_index=255))
    ref.backward()
    assert rel_err(loss['loss_ce'], ref) <= LOSS_TOL
    assert rel_err(x.grad, xo.grad) <= GRAD_TOL
    # fused entry: every scalar it hands out accepts an in-place update as well
    r = B.fused_resize_losses(x, y.unsqueeze(1), [B.CrossEntropyLoss(), B.DiceLoss()], ignore_index=255)
    r['loss_ce'] += r['loss_dice']
    r['acc_seg'] += 1.0
    r['loss_ce'].backward()


def test_bce_elementwise_forwards_scalar_class_weight(B):
    """cross_entropy_loss.py:160-161: pos_weight=class_weight on every branch, element-wise targets included."""
    g = torch.Generator().manual_seed(5)
    p = torch.randn((4, 3, 8, 8), generator=g).cuda().requires_grad_(True)
    t = (torch.rand((4, 3, 8, 8), generator=g) > 0.5).float().cuda()
    a = B.binary_cross_entropy(p, t, class_weight=[2.5])
    a.backward()
    po = p.detach().clone().requires_grad_(True)
    b = O.binary_cross_entropy(po, t, class_weight=torch.tensor([2.5], device='cuda'))
    b.backward()
    assert rel_err(a, b) <= LOSS_TOL and rel_err(p.grad, po.grad) <= GRAD_TOL
    with pytest.raises(NotImplementedError):
        B.binary_cross_entropy(p, t, class_weight=[1.0, 2.0, 3.0])
    h = B.binary_cross_entropy(p.detach().bfloat16(), t, class_weight=[2.5])
    assert h.dtype == torch.bfloat16


def test_second_device_in_one_process(B):
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per (kernel, device): the first launch on a second GPU of the same
    process must opt in again (confusion kernels with C >= 17, the bulk-copy pipeline, the cell-owner kernel)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 visible GPUs')
    res = []
    for dev in ('cuda:0', 'cuda:1'):
        x = synth_logits((2, 21, 64, 96), 2, device=dev).requires_grad_(True)
        y = synth_labels((2, 64, 96), 21, 2, device=dev, block=8)
        l = B.CrossEntropyLoss()(x, y, ignore_index=255)
        l.backward()
        xl = synth_logits((2, 19, 8, 16), 4, device=dev).requires_grad_(True)
        yl = synth_labels((2, 64, 128), 19, 4, device=dev, block=8)
        r = B.fused_resize_losses(xl, yl.unsqueeze(1), B.CrossEntropyLoss(), ignore_index=255)
        r['loss_ce'].backward()
        preds = [torch.randint(0, 19, (64, 80), device=dev) for _ in range(3)]
        gts = [torch.randint(0, 19, (64, 80), device=dev).float() for _ in range(3)]
        areas = B.areas_device(preds, gts, 19, 255)
        res.append((float(l), x.grad.cpu(), float(r['loss_ce']), xl.grad.cpu(), areas.cpu()))
    torch.cuda.synchronize()
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1])
    assert res[0][2] == res[1][2] and torch.equal(res[0][3], res[1][3])
    assert int(res[0][4].sum()) > 0


def test_multi_gpu_nccl_sharded_totals_match_single_gpu(B, tmp_path):
    """Row (e): two NCCL ranks shard a config-4-shaped batch and a 16-image config-5 list with shard_range; the
    all-reduced loss scalars / int64 areas must equal the single-GPU result (areas bit-for-bit, loss to 1e-6)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 visible GPUs')
    out = tmp_path / 'mgpu.pt'
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''), MGPU_OUT=str(out))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', '29633', os.path.join(ROOT, 'tests', 'mgpu_worker.py')]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    got = torch.load(str(out))
    # single GPU, whole batch
    N, C = 8, 21
    x = synth_logits((N, C, 128, 128), 77, device='cuda').requires_grad_(True)
    y = synth_labels((N, 128, 128), C, 77, device='cuda', block=8)
    rs = B.fused_resize_losses(x, y.unsqueeze(1), [B.CrossEntropyLoss(), B.DiceLoss(loss_weight=3.0)], ignore_index=255,
                               return_stats=True)
    (rs['loss_ce'] + rs['loss_dice']).backward()
    assert abs(got['loss_ce'] - float(rs['loss_ce'])) <= 1e-6 * abs(float(rs['loss_ce']))
    assert abs(got['loss_dice'] - float(rs['loss_dice'])) <= 1e-6 * abs(float(rs['loss_dice']))
    assert abs(got['acc_seg'] - float(rs['acc_seg'])) <= 1e-4
    assert rel_err(got['grad'], x.grad) <= 1e-6
    preds = [torch.randint(0, 19, (96, 160), generator=torch.Generator().manual_seed(900 + i)).cuda() for i in range(16)]
    gts = [synth_labels((1, 96, 160), 19, 900 + i, device='cuda', block=8)[0].float() for i in range(16)]
    tot = B.areas_device(preds, gts, 19, 255).sum(0).cpu()
    assert got['areas'].dtype == torch.int64 and torch.equal(got['areas'], tot)
    # parse_losses: rank mean of the per-rank 'mean' losses == the global-batch loss for equal shards
    assert abs(got['logged']['loss_ce'] - float(rs['loss_ce'])) <= 1e-5 * abs(float(rs['loss_ce']))
    assert abs(got['logged']['loss'] - float(rs['loss_ce'] + rs['loss_dice'])) <= 1e-5 * abs(float(rs['loss_ce'] + rs['loss_dice']))


# ------------------------------------------------------------------------------------------------ class-sliced pipeline
def _ce_dice(B, x, y, ce_kw, dice_kw, pw=None, scale=1.0):
    xa = x.clone().requires_grad_(True)
    r = B.fused_resize_losses(xa, y.unsqueeze(1), [B.CrossEntropyLoss(**ce_kw), B.DiceLoss(**dice_kw)], ignore_index=255,
                              seg_weight=pw)
    ((r['loss_ce'] + r['loss_dice']) * scale).backward()
    return r['loss_ce'].detach(), r['loss_dice'].detach(), r['acc_seg'].detach(), xa.grad


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16])
def test_class_sliced_pipeline_matches_oracle_and_streaming_kernels(B, dtype):
    """csrc/loss_cs.cu (one read of the logits per direction, 32 < C <= 152) against the oracle and against the five-pass
    streaming kernels it supersedes (B200SEG_NO_CS=1), over every class-slice width, ragged tiles, label dtypes, pixel
    weights, a skipped Dice class and an upstream gradient."""
    cases = [((2, 150, 64, 64), {}, dict(loss_weight=3.0), False, torch.int64, 1.0),
             ((3, 40, 24, 40), dict(class_weight=torch.linspace(0.5, 1.5, 40).tolist()), dict(), True, torch.uint8, 1.0),
             ((1, 100, 32, 32), dict(avg_non_ignore=True), dict(ignore_index=3, class_weight=[0.7] * 100), False, torch.int32, 8.0),
             ((2, 70, 16, 24), dict(reduction='sum'), dict(smooth=2.0), False, torch.float32, 1.0),
             ((2, 150, 24, 40), dict(class_weight=torch.linspace(0.5, 1.5, 150).tolist()), dict(loss_weight=3.0), True, torch.int64, 1.0),
             ((1, 33, 8, 16), {}, dict(), False, torch.int64, 1.0),
             # H*W = 136 / 168: the last tile's TMA box is only partly inside the image (out-of-range columns read as 0)
             ((2, 40, 8, 17), {}, dict(), False, torch.int64, 1.0),
             ((1, 150, 12, 14), dict(class_weight=torch.linspace(0.5, 1.5, 150).tolist()), dict(loss_weight=3.0), True, torch.int64, 1.0)]
    tol_l, tol_g = (LOSS_TOL, GRAD_TOL) if dtype == torch.float32 else (HALF_TOL, 2 * HALF_TOL)
    for shape, ce_kw, dice_kw, with_pw, ldt, scale in cases:
        n, c, h, w = shape
        x = synth_logits(shape, 31, dtype=dtype, device='cuda')
        y = synth_labels((n, h, w), c, 31, device='cuda', block=4).to(ldt)
        pw = (torch.rand((n, h, w), device='cuda') + 0.5) if with_pw else None
        got = _ce_dice(B, x, y, ce_kw, dice_kw, pw, scale)
        os.environ['B200SEG_NO_CS'] = '1'
        try:
            old = _ce_dice(B, x, y, ce_kw, dice_kw, pw, scale)
        finally:
            del os.environ['B200SEG_NO_CS']
        xo = x.float().requires_grad_(True)
        ce_o = dict(ce_kw)
        dice_o = dict(dice_kw)
        for kw in (ce_o, dice_o):
            if 'class_weight' in kw:
                kw['class_weight'] = list(kw['class_weight'])
        lce = O.cross_entropy_loss_module(xo, y.long(), weight=pw, ignore_index=255, **ce_o)
        ldi = O.dice_loss_module(xo, y.long(), **dice_o)
        ((lce + ldi) * scale).backward()
        name = '%s %s' % (shape, dtype)
        assert rel_err(got[0], lce) <= tol_l, name
        assert rel_err(got[1], ldi) <= tol_l or float((got[1].double().cpu() - ldi.double().cpu()).abs()) < 2e-6, name
        assert rel_err(got[3], xo.grad) <= tol_g, (name, rel_err(got[3], xo.grad))
        # the two CUDA paths agree far below the gate (same fp32 arithmetic, different summation order)
        assert rel_err(got[0], old[0]) <= 1e-6 and rel_err(got[1], old[1]) <= (1e-5 if dtype == torch.float32 else 1e-3), name
        assert rel_err(got[3].float(), old[3].float()) <= (2e-5 if dtype == torch.float32 else 2.0 ** -7), name
        if dtype == torch.float32:
            assert abs(float(got[2]) - float(O.accuracy(x, y.long(), ignore_index=255))) <= 1e-3, name
    # forward only (no_grad), and determinism of the gradient
    x = synth_logits((2, 150, 32, 64), 5, dtype=dtype, device='cuda')
    y = synth_labels((2, 32, 64), 150, 5, device='cuda', block=4)
    with torch.no_grad():
        r = B.fused_resize_losses(x, y.unsqueeze(1), [B.CrossEntropyLoss(), B.DiceLoss()], ignore_index=255)
    a = _ce_dice(B, x, y, {}, {})
    b = _ce_dice(B, x, y, {}, {})
    assert rel_err(r['loss_ce'], a[0]) <= 1e-6
    assert torch.equal(a[3], b[3]) or rel_err(a[3].float(), b[3].float()) <= 1e-6   # dice sums are fp64 atomics


# ------------------------------------------------------------------------------------------------ thread-per-cell resize-fused CE
def _up_case(B, shape, size, C, ac, dtype=torch.float32, ce_kw=None, pixel_weight=False, ldt=torch.int64, seed=3, scale=1.0):
    n = shape[0]
    x = (synth_logits(shape, seed, dtype=torch.float32, device='cuda') * scale).to(dtype)
    y = synth_labels((n,) + tuple(size), C, seed, device='cuda', block=5).to(ldt)
    pw = (torch.rand((n,) + tuple(size), device='cuda') + 0.5) if pixel_weight else None
    ce_kw = ce_kw or {}
    xa = x.clone().requires_grad_(True)
    r = B.fused_resize_losses(xa, y.unsqueeze(1), B.CrossEntropyLoss(**ce_kw), align_corners=ac, ignore_index=255, seg_weight=pw)
    r['loss_ce'].backward()
    xo = x.float().requires_grad_(True)
    full = O.resize(xo, size=size, mode='bilinear', align_corners=ac)
    kw = dict(ce_kw)
    lo = O.cross_entropy_loss_module(full, y.long(), weight=pw, ignore_index=255, **kw)
    lo.backward()
    acc = O.accuracy(full.detach(), y.long(), ignore_index=255)
    return (r['loss_ce'].detach(), xa.grad, r['acc_seg']), (lo.detach(), xo.grad, acc)


@pytest.mark.parametrize('ac', [False, True])
def test_resize_fused_any_ratio_single_pass(B, ac):
    """csrc/loss_upgen.cuh: the resize-fused single pass for ANY up-sampling ratio and both align_corners settings
    (utils/ops.py:7-26 with align_corners=self.align_corners, decode_head.py:266-269): odd ratios, different ratios per
    axis, one axis at label resolution, tiny extents, every class count up to 32, weights, label dtypes, 16-bit logits."""
    from image_segmentation_lab_b200 import _lib
    lib = _lib.load()
    cases = [((2, 19, 65, 129), (513, 1025), 19, {}, False, torch.int64, torch.float32),
             ((2, 19, 64, 128), (512, 1024), 19, {}, False, torch.int64, torch.float32),
             ((1, 7, 10, 13), (37, 91), 7, dict(class_weight=[0.5 + 0.1 * i for i in range(7)]), True, torch.uint8, torch.float32),
             ((2, 32, 9, 11), (20, 50), 32, dict(avg_non_ignore=True), False, torch.int32, torch.float32),   # ratio ~2.2 / 4.5
             ((2, 5, 16, 12), (16, 96), 5, dict(reduction='sum'), False, torch.int64, torch.float32),        # H == h
             ((1, 3, 1, 1), (9, 7), 3, {}, False, torch.int64, torch.float32),
             ((2, 21, 17, 23), (130, 180), 21, {}, True, torch.float32, torch.float32),
             ((2, 19, 33, 65), (257, 513), 19, {}, False, torch.int64, torch.bfloat16),
             ((2, 19, 16, 32), (256, 512), 19, {}, False, torch.int64, torch.float16)]
    for shape, size, C, kw, pwt, ldt, dtype in cases:
        n, c, h, w = shape
        assert lib.b200seg_loss_fused_workspace_bytes(n, c, h, w, size[0], size[1], int(ac)) > 0
        got, ref = _up_case(B, shape, size, C, ac, dtype=dtype, ce_kw=kw, pixel_weight=pwt, ldt=ldt)
        tl, tg = (LOSS_TOL, GRAD_TOL) if dtype == torch.float32 else (HALF_TOL, 2 * HALF_TOL)
        name = '%s->%s ac=%s %s' % (shape, size, ac, dtype)
        assert rel_err(got[0], ref[0]) <= tl, (name, rel_err(got[0], ref[0]))
        assert rel_err(got[1], ref[1]) <= tg, (name, rel_err(got[1], ref[1]))
        # exact ties between interpolated classes count as correct for the label here, torch.topk picks by its own rule:
        # allow two pixels on the small shapes
        npx = n * size[0] * size[1]
        assert abs(float(got[2]) - float(ref[2])) <= (max(0.02, 200.0 / npx) if dtype == torch.float32 else 0.5), name
    # steep logits: classes hundreds of nats below the cell maximum -> the direct-evaluation fallback
    for sc in (20.0, 90.0):
        got, ref = _up_case(B, (2, 19, 9, 13), (70, 100), 19, ac, scale=sc)
        assert rel_err(got[0], ref[0]) <= 2e-5 and rel_err(got[1], ref[1]) <= GRAD_TOL, sc


def test_resize_fused_paths_of_the_packed_kernel(B):
    """Code paths the packed-math kernel added (csrc/loss_upgen.cuh): class counts that need 0..3 pad classes, every row-group
    count (1, 2, 4 threads per cell), labels outside [0, C) counted in the statistics and treated as ignored, long runs
    of ignored pixels inside a chunk (the run-merged one-hot update), uint8 labels at all four alignments of a chunk, and
    horizontal logit differences steep enough for R^(PXC/2) to overflow (cell-level fallback)."""
    from image_segmentation_lab_b200 import _lib
    # class counts: pads 3, 2, 1, 0; tiny class counts
    for C in (1, 2, 3, 4, 17, 18, 20, 31):
        got, ref = _up_case(B, (2, C, 12, 20), (96, 160), C, False, seed=C)
        if C == 1:      # the reference's loss and gradient are exactly 0: absolute bounds (the chain rounds at 2^-23)
            assert float(got[0].abs()) <= 1e-6 and float(got[1].abs().max()) <= 1e-9
            continue
        assert rel_err(got[0], ref[0]) <= LOSS_TOL and rel_err(got[1], ref[1]) <= GRAD_TOL, C
    # row groups: scale 2 (1 thread per cell), scale 4 (2), scale 8 and 16 (4); a single cell column / row
    for shape, size in (((1, 19, 40, 50), (80, 100)), ((1, 19, 40, 50), (160, 200)), ((1, 19, 20, 24), (160, 192)),
                        ((1, 19, 6, 8), (96, 128)), ((1, 5, 1, 9), (16, 72)), ((1, 5, 9, 1), (72, 16))):
        for ac in (False, True):
            got, ref = _up_case(B, shape, size, shape[1], ac, seed=11)
            assert rel_err(got[0], ref[0]) <= LOSS_TOL and rel_err(got[1], ref[1]) <= GRAD_TOL, (shape, size, ac)
    # uint8 labels whose rows start at every alignment (W = 157: rows shift by 1 byte), int64 rows at odd offsets (W odd)
    for ldt in (torch.uint8, torch.int64):
        got, ref = _up_case(B, (2, 19, 10, 20), (80, 157), 19, False, ldt=ldt, seed=5)
        assert rel_err(got[0], ref[0]) <= LOSS_TOL and rel_err(got[1], ref[1]) <= GRAD_TOL, ldt
    # labels outside [0, C): counted, and ignored by the loss, its gradient and the accuracy numerator
    x = synth_logits((2, 19, 16, 32), 21, device='cuda')
    y = synth_labels((2, 128, 256), 19, 21, device='cuda')
    y_bad = y.clone()
    y_bad[:, 10:30, 40:44] = 77
    y_bad[1, 100, :] = -3
    n_bad = int(((y_bad != 255) & ((y_bad < 0) | (y_bad >= 19))).sum())
    xa = x.clone().requires_grad_(True)
    r = B.fused_resize_losses(xa, y_bad.unsqueeze(1), B.CrossEntropyLoss(), ignore_index=255, return_stats=True)
    r['loss_ce'].backward()
    assert int(r['_stats'][_lib.LOG_N_BAD]) == n_bad and n_bad > 0
    y_ign = torch.where((y_bad < 0) | ((y_bad >= 19) & (y_bad != 255)), torch.full_like(y_bad, 255), y_bad)
    xo = x.clone().requires_grad_(True)
    full = O.resize(xo, size=(128, 256), mode='bilinear', align_corners=False)
    lo = O.cross_entropy_loss_module(full, y_ign, ignore_index=255)
    lo.backward()
    assert rel_err(r['loss_ce'].detach(), lo.detach()) <= LOSS_TOL and rel_err(xa.grad, xo.grad) <= GRAD_TOL
    # a label map that is mostly ignore_index, with isolated valid pixels (runs of length 1 between ignored stretches)
    y_sparse = torch.full_like(y, 255)
    y_sparse[:, ::3, ::5] = y[:, ::3, ::5]
    xa = x.clone().requires_grad_(True)
    r = B.fused_resize_losses(xa, y_sparse.unsqueeze(1), B.CrossEntropyLoss(avg_non_ignore=True), ignore_index=255)
    r['loss_ce'].backward()
    xo = x.clone().requires_grad_(True)
    lo = O.cross_entropy_loss_module(O.resize(xo, size=(128, 256), mode='bilinear', align_corners=False), y_sparse, ignore_index=255,
                                     avg_non_ignore=True)
    lo.backward()
    assert rel_err(r['loss_ce'].detach(), lo.detach()) <= LOSS_TOL and rel_err(xa.grad, xo.grad) <= GRAD_TOL
    # neighbouring logits 60 nats apart along x at scale ~1.5: R^(PXC/2) leaves the normal range -> cell-level fallback
    xs = synth_logits((1, 6, 8, 40), 4, device='cuda')
    xs[:, :, :, ::2] += 60.0
    ys = synth_labels((1, 12, 60), 6, 4, device='cuda')
    xa = xs.clone().requires_grad_(True)
    r = B.fused_resize_losses(xa, ys.unsqueeze(1), B.CrossEntropyLoss(), ignore_index=255)
    r['loss_ce'].backward()
    xo = xs.clone().requires_grad_(True)
    lo = O.cross_entropy_loss_module(O.resize(xo, size=(12, 60), mode='bilinear', align_corners=False), ys, ignore_index=255)
    lo.backward()
    assert torch.isfinite(r['loss_ce']) and rel_err(r['loss_ce'].detach(), lo.detach()) <= 2e-5 and rel_err(xa.grad, xo.grad) <= GRAD_TOL


def test_resize_fused_general_is_deterministic_and_matches_round1_kernel(B):
    """No atomics anywhere in the resize-fused backward: bitwise identical gradients run to run for align_corners=True and
    odd ratios; at power-of-two ratios the thread-per-cell kernel agrees with the round-1 quad-per-cell kernel."""
    for shape, size, ac in (((3, 19, 33, 65), (257, 513), True), ((2, 19, 30, 40), (97, 131), False)):
        x = synth_logits(shape, 8, device='cuda')
        y = synth_labels((shape[0],) + size, 19, 8, device='cuda').unsqueeze(1)
        grads = []
        for _ in range(3):
            xa = x.clone().requires_grad_(True)
            B.fused_resize_losses(xa, y, B.CrossEntropyLoss(), align_corners=ac, ignore_index=255)['loss_ce'].backward()
            grads.append(xa.grad.clone())
        assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])
    x = synth_logits((4, 19, 32, 64), 8, device='cuda')
    y = synth_labels((4, 256, 512), 19, 8, device='cuda').unsqueeze(1)
    res = []
    for old in ('gen', 'old'):
        os.environ['B200SEG_UPCELL'] = old
        try:
            xa = x.clone().requires_grad_(True)
            r = B.fused_resize_losses(xa, y, B.CrossEntropyLoss(), ignore_index=255)
            r['loss_ce'].backward()
            res.append((r['loss_ce'].detach().clone(), xa.grad.clone(), r['acc_seg'].clone()))
        finally:
            del os.environ['B200SEG_UPCELL']
    assert rel_err(res[0][0], res[1][0]) <= 2e-6 and rel_err(res[0][1], res[1][1]) <= 1e-5
    assert abs(float(res[0][2]) - float(res[1][2])) <= 1e-3


def test_resize_scale_factor_and_nearest_backward(B):
    """utils/ops.py:7-26 passes size= / scale_factor= and any mode through to F.interpolate: fractional scale factors (the
    source-index scale becomes 1 / scale_factor, not in / out) and the nearest mode with a gradient."""
    F = torch.nn.functional
    x = torch.randn(2, 3, 11, 14, device='cuda')
    for sf in (2, 1.5, (2.5, 1.7), 0.6):
        for ac in (False, True):
            xa = x.clone().requires_grad_(True)
            xb = x.clone().requires_grad_(True)
            ya = B.resize(xa, scale_factor=sf, mode='bilinear', align_corners=ac, warning=False)
            yb = F.interpolate(xb, scale_factor=sf, mode='bilinear', align_corners=ac)
            assert ya.shape == yb.shape and torch.equal(ya, yb), (sf, ac)
            go = torch.randn_like(yb)
            ya.backward(go)
            yb.backward(go)
            assert rel_err(xa.grad, xb.grad) <= 1e-5, (sf, ac)
        xa = x.clone().requires_grad_(True)
        xb = x.clone().requires_grad_(True)
        ya = B.resize(xa, scale_factor=sf)              # default mode: nearest
        yb = F.interpolate(xb, scale_factor=sf)
        assert torch.equal(ya, yb), sf
        go = torch.randn_like(yb)
        ya.backward(go)
        yb.backward(go)
        assert rel_err(xa.grad, xb.grad) <= 1e-5, sf
    y = B.resize(x, size=(23, 9))
    assert torch.equal(y, F.interpolate(x, size=(23, 9)))
    with pytest.raises(NotImplementedError):
        B.resize(x, size=(20, 20), mode='bicubic')


def test_parse_losses_on_device_step(B):
    """parse_losses (utils/train_utils.py:31-74) over the dict the fused head returns, on the GPU: the summed
    loss stays in the autograd graph, the logged values equal the reference's per-variable .item() reads, and the lazy
    form does not synchronise (it is capturable in a CUDA graph together with the step)."""
    x = synth_logits((2, 19, 16, 32), 5, device='cuda').requires_grad_(True)
    y = synth_labels((2, 128, 256), 19, 5, device='cuda').unsqueeze(1)
    losses = B.fused_resize_losses(x, y, [B.CrossEntropyLoss(), B.DiceLoss(loss_weight=3.0)], ignore_index=255, return_stats=True)
    loss, log_vars = B.parse_losses(losses)
    loss.backward()
    xo = x.detach().clone().requires_grad_(True)
    ref = O.head_losses(xo, y, [('ce', {}, 'loss_ce'), ('dice', dict(loss_weight=3.0), 'loss_dice')], ignore_index=255)
    ref_loss = ref['loss_ce'] + ref['loss_dice']
    ref_loss.backward()
    assert list(log_vars.keys()) == ['loss_ce', 'loss_dice', 'acc_seg', 'loss']
    assert abs(log_vars['loss'] - float(ref_loss)) <= LOSS_TOL * abs(float(ref_loss))
    assert abs(log_vars['loss_ce'] - float(ref['loss_ce'])) <= LOSS_TOL * abs(float(ref['loss_ce']))
    assert abs(log_vars['acc_seg'] - float(ref['acc_seg'])) <= 1e-3
    assert rel_err(x.grad, xo.grad) <= GRAD_TOL
    # lazy: device scalars only — the whole step + parse_losses captures into one CUDA graph
    xs = x.detach().clone().requires_grad_(True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        B.parse_losses(B.fused_resize_losses(xs, y, B.CrossEntropyLoss(), ignore_index=255), lazy=True)[0].backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    xs.grad = None
    with torch.cuda.graph(g):
        l2, lazy = B.parse_losses(B.fused_resize_losses(xs, y, B.CrossEntropyLoss(), ignore_index=255), lazy=True)
        l2.backward()
    g.replay()
    torch.cuda.synchronize()
    r1 = O.head_losses(x.detach(), y, [('ce', {}, 'loss_ce')], ignore_index=255)
    assert abs(float(lazy['loss']) - float(r1['loss_ce'])) <= LOSS_TOL * abs(float(r1['loss_ce']))


def test_bf16_against_the_native_bf16_oracle(B):
    """BASELINE.md section 5 / SURVEY 8d list two bf16 gates: the fp32-upcast oracle at 2**-7 (test_gpu_parity.py) AND the
    reference run natively in bf16 (what it does outside autocast) at bf16 tolerance (~1e-2): every intermediate of the
    reference is rounded to bf16 there, this path computes in fp32 and rounds once."""
    cw = torch.linspace(0.5, 1.5, 150).tolist()
    for shape, C, ce_kw, dice_kw in (((2, 150, 64, 64), 150, dict(class_weight=cw), dict(loss_weight=3.0)),
                                     ((2, 19, 64, 64), 19, {}, dict()), ((2, 21, 32, 48), 21, {}, None)):
        x = synth_logits(shape, 17, dtype=torch.bfloat16, device='cuda')
        y = synth_labels((shape[0],) + shape[2:], C, 17, device='cuda', block=8)
        xa = x.clone().requires_grad_(True)
        mods = [B.CrossEntropyLoss(**ce_kw)] + ([B.DiceLoss(**dice_kw)] if dice_kw is not None else [])
        r = B.fused_resize_losses(xa, y.unsqueeze(1), mods, ignore_index=255)
        tot = r['loss_ce'] + (r['loss_dice'] if dice_kw is not None else 0)
        tot.backward()
        xb = x.clone().requires_grad_(True)          # native bf16: the oracle's ATen calls run in bf16
        lce = O.cross_entropy_loss_module(xb, y, ignore_index=255, **ce_kw)
        ldi = O.dice_loss_module(xb, y, **dice_kw) if dice_kw is not None else None
        assert lce.dtype == torch.bfloat16
        (lce + (ldi if ldi is not None else 0)).backward()
        assert rel_err(r['loss_ce'], lce) <= 1e-2, shape
        if ldi is not None:
            assert rel_err(r['loss_dice'], ldi) <= 5e-2, shape     # the reference's bf16 running sums over 150 classes
        assert rel_err(xa.grad.float(), xb.grad.float()) <= 3e-2, (shape, rel_err(xa.grad.float(), xb.grad.float()))


def test_resize_fused_many_classes_class_tiled(B):
    """C > 32 with low-resolution logits (ADE20K's 150 classes at 1/8 resolution): the class-tiled plan of
    csrc/loss_upgen.cuh — one forward launch over all classes, one backward launch per tile of 32 classes — against the
    oracle, deterministic, without materialising the (N,C,H,W) tensor."""
    from image_segmentation_lab_b200 import _lib
    lib = _lib.load()
    for shape, size, ac, C, kw, dtype in (((2, 150, 16, 16), (128, 128), False, 150, {}, torch.float32),
                                          ((1, 60, 9, 13), (70, 100), True, 60, dict(class_weight=[0.5 + 0.01 * i for i in range(60)]), torch.float32),
                                          ((2, 33, 8, 8), (32, 32), False, 33, dict(avg_non_ignore=True), torch.float32),
                                          ((2, 150, 16, 16), (128, 128), False, 150, {}, torch.bfloat16)):
        n, c, h, w = shape
        assert lib.b200seg_loss_fused_workspace_bytes(n, c, h, w, size[0], size[1], int(ac)) == \
            n * c * (h + 1) * (w + 1) * 16 + n * size[0] * size[1] * 4
        got, ref = _up_case(B, shape, size, C, ac, dtype=dtype, ce_kw=kw)
        tl, tg = (LOSS_TOL, GRAD_TOL) if dtype == torch.float32 else (HALF_TOL, 2 * HALF_TOL)
        assert rel_err(got[0], ref[0]) <= tl, (shape, rel_err(got[0], ref[0]))
        assert rel_err(got[1], ref[1]) <= tg, (shape, rel_err(got[1], ref[1]))
        assert abs(float(got[2]) - float(ref[2])) <= (max(0.02, 200.0 / (n * size[0] * size[1])) if dtype == torch.float32 else 0.5)
    # steep logits -> direct evaluation inside the tiles; and bitwise determinism
    got, ref = _up_case(B, (2, 40, 9, 13), (70, 100), 40, False, scale=60.0)
    assert rel_err(got[0], ref[0]) <= 2e-5 and rel_err(got[1], ref[1]) <= GRAD_TOL
    x = synth_logits((2, 150, 16, 16), 8, device='cuda')
    y = synth_labels((2, 128, 128), 150, 8, device='cuda').unsqueeze(1)
    grads = []
    for _ in range(2):
        xa = x.clone().requires_grad_(True)
        B.fused_resize_losses(xa, y, B.CrossEntropyLoss(), ignore_index=255)['loss_ce'].backward()
        grads.append(xa.grad.clone())
    assert torch.equal(grads[0], grads[1])
    with torch.no_grad():
        r = B.fused_resize_losses(x, y, B.CrossEntropyLoss(), ignore_index=255)
    ro = O.head_losses(x, y, [('ce', {}, 'loss_ce')], ignore_index=255)
    assert rel_err(r['loss_ce'], ro['loss_ce']) <= LOSS_TOL


def test_bce_single_pass_matches_two_pass_and_oracle(B):
    """Sigmoid CE (row f1): the single-pass plan (gradient written by the forward launch, reduced scalar by its last CTA)
    against the two-pass plan and the oracle — loss_weight, pixel weights, pos_weight, every reduction with a known
    denominator, an upstream gradient != 1 (the late rescale), bf16; float16 and avg_non_ignore stay two-pass."""
    x0 = synth_logits((3, 4, 40, 56), 17, device='cuda', margin=False)
    y = synth_labels((3, 40, 56), 4, 17, ignore_index=255, block=4, device='cuda')
    w = torch.rand((3, 40, 56), device='cuda') + 0.5
    for kw, fkw in ((dict(), dict()), (dict(reduction='sum', loss_weight=0.3), dict()),
                    (dict(loss_weight=2.0), dict(weight=w)),
                    (dict(), dict(avg_factor=1234.0)), (dict(avg_non_ignore=True), dict())):
        res = []
        for single in (True, False):
            m = B.CrossEntropyLoss(use_sigmoid=True, **kw)
            m.single_pass = single
            x = x0.clone().requires_grad_(True)
            loss = m(x, y, ignore_index=255, **fkw)
            (loss * 3.0).backward()
            res.append((loss.detach(), x.grad))
        xo = x0.clone().requires_grad_(True)
        kwo = dict(kw)
        lw = kwo.pop('loss_weight', 1.0)
        if 'class_weight' in kwo:
            kwo['class_weight'] = xo.new_tensor(kwo['class_weight'])
        lo = lw * O.binary_cross_entropy(xo, y, fkw.get('weight'), ignore_index=255, avg_factor=fkw.get('avg_factor'), **kwo)
        (lo * 3.0).backward()
        for loss, grad in res:
            assert rel_err(loss, lo) <= LOSS_TOL, (kw, float(loss), float(lo))
            assert rel_err(grad, xo.grad) <= GRAD_TOL, kw
    # pos_weight = class_weight (:160-161) on (N,C) predictions, where the reference's broadcast lines up with the classes
    g = torch.Generator().manual_seed(3)
    p2 = torch.randn((600, 5), generator=g).cuda()
    y2 = torch.randint(0, 5, (600,), generator=g).cuda()
    y2[::7] = 255
    cwl = [0.5, 1.0, 2.0, 1.5, 3.0]
    for single in (True, False):
        m = B.CrossEntropyLoss(use_sigmoid=True, class_weight=cwl, loss_weight=1.7)
        m.single_pass = single
        x = p2.clone().requires_grad_(True)
        m(x, y2, ignore_index=255).backward()
        xo = p2.clone().requires_grad_(True)
        lo = 1.7 * O.binary_cross_entropy(xo, y2, ignore_index=255, class_weight=xo.new_tensor(cwl))
        lo.backward()
        assert rel_err(x.grad, xo.grad) <= GRAD_TOL
    # the head call: a lone sigmoid loss also yields the top-1 accuracy from the same launch (decode_head.py:295)
    x = x0.clone().requires_grad_(True)
    r = B.fused_resize_losses(x, y.unsqueeze(1), B.CrossEntropyLoss(use_sigmoid=True), ignore_index=255)
    r['loss_ce'].backward()
    xo = x0.clone().requires_grad_(True)
    lo = O.binary_cross_entropy(xo, y, ignore_index=255)
    lo.backward()
    assert rel_err(r['loss_ce'], lo) <= LOSS_TOL and rel_err(x.grad, xo.grad) <= GRAD_TOL
    ao = O.accuracy(x0, y, ignore_index=255)
    assert r['acc_seg'].shape == (1,) and abs(float(r['acc_seg']) - float(ao)) <= 1e-4 * float(ao)
    # the single-pass graph is consumed by its backward (like the soft-max path's flat plan)
    m = B.CrossEntropyLoss(use_sigmoid=True)
    x = x0.clone().requires_grad_(True)
    loss = m(x, y, ignore_index=255)
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError):
        loss.backward()
    # float16: two-pass (a large upstream gradient must meet the 1/numel factor in fp32), as ADVICE r1 asked of the flat plan
    xh = x0.half().requires_grad_(True)
    lh = B.CrossEntropyLoss(use_sigmoid=True)(xh, y, ignore_index=255)
    (lh.float() * 16384.0).backward()                     # (the loss itself comes back as float16: 65536 would overflow it)
    xr = x0.half().float().requires_grad_(True)
    (O.binary_cross_entropy(xr, y, ignore_index=255) * 16384.0).backward()
    assert rel_err(xh.grad.float(), xr.grad) <= 2e-3
    # no_grad / forward only
    with torch.no_grad():
        ln = B.CrossEntropyLoss(use_sigmoid=True)(x0, y, ignore_index=255)
    assert rel_err(ln, O.binary_cross_entropy(x0, y, ignore_index=255)) <= LOSS_TOL


# ------------------------------------------------------------------------------------------------ forward-only bulk pipeline
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16])
def test_bulk_forward_only_matches_oracle_and_streaming_forward(B, dtype):
    """csrc/loss_bulk.cu in its forward-only form (validation loss under no_grad, reduction='none', the first pass of the
    two-pass plans: cross_entropy_loss.py:56-72) against the oracle and against the streaming forward it replaces
    (B200SEG_NO_BULK_FWD=1): reductions, class weights, avg_non_ignore + backward through the saved log-sum-exp, label
    dtypes, a ragged last tile, out-of-range labels, all-ignored input."""
    tol = LOSS_TOL if dtype == torch.float32 else HALF_TOL
    cases = [((4, 21, 64, 64), {}, torch.int64), ((2, 19, 24, 40), dict(reduction='sum'), torch.uint8),
             ((2, 21, 20, 28), dict(reduction='none'), torch.int64),
             ((3, 8, 16, 24), dict(avg_non_ignore=True, class_weight=torch.linspace(0.5, 1.5, 8).tolist()), torch.int32),
             ((1, 34 if dtype == torch.float32 else 69, 32, 36), dict(class_weight=None), torch.int64), ((2, 2, 8, 8), {}, torch.float32)]
    for shape, kw, ldt in cases:
        n, c, h, w = shape
        x = synth_logits(shape, 77, dtype=dtype, device='cuda')
        y = synth_labels((n, h, w), c, 77, device='cuda', block=4)
        yo = y.clone()
        if not kw.get('avg_non_ignore'):         # (a bad label counts in avg_non_ignore's denominator; ATen device-asserts on it)
            y[0, 0, :3] = c + 5                  # out of range, not ignore_index: counted as bad, treated as ignored
            yo[0, 0, :3] = 255
        y = y.to(ldt)
        res = {}
        for mode in ('bulk', 'stream'):
            if mode == 'stream':
                os.environ['B200SEG_NO_BULK_FWD'] = '1'
            try:
                with torch.no_grad():
                    l = B.CrossEntropyLoss(**kw)(x, y, ignore_index=255)
                    a = B.accuracy(x, yo, ignore_index=255)
                xa = x.clone().requires_grad_(True)
                ce = B.CrossEntropyLoss(**kw)
                ce.single_pass = False           # two-pass plan: forward saves the log-sum-exp, backward re-reads the logits
                lg = ce(xa, y, ignore_index=255)
                lg.sum().backward()
                res[mode] = (l, a, lg.detach(), xa.grad)
            finally:
                os.environ.pop('B200SEG_NO_BULK_FWD', None)
        xo = x.float().requires_grad_(True)
        want = O.cross_entropy_loss_module(xo, yo.long(), ignore_index=255, **kw)
        want.sum().backward()
        name = '%s %s %s' % (shape, kw, dtype)
        for mode in ('bulk', 'stream'):
            l, a, lg, g = res[mode]
            assert l.shape == want.shape, name
            assert rel_err(l, want) <= tol and rel_err(lg, want) <= tol, (name, mode, rel_err(l, want))
            assert rel_err(g, xo.grad) <= (GRAD_TOL if dtype == torch.float32 else 2 * HALF_TOL), (name, mode)
        assert rel_err(res['bulk'][0], res['stream'][0]) <= 2e-6, name
        assert torch.equal(res['bulk'][1], res['stream'][1]), name          # top-1 accuracy: integer counts
        assert rel_err(res['bulk'][3].float(), res['stream'][3].float()) <= (2e-5 if dtype == torch.float32 else 2.0 ** -7), name
    x = synth_logits((2, 21, 16, 16), 3, dtype=dtype, device='cuda')
    y = torch.full((2, 16, 16), 255, device='cuda')
    with torch.no_grad():
        assert float(B.CrossEntropyLoss()(x, y, ignore_index=255)) == 0.0
