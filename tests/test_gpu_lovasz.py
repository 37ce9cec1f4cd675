"""GPU: LovaszLoss (models/losses/lovasz_loss.py:26-312) on csrc/loss_lovasz.cu, through the Python mirror -> C ABI.

Gates: loss <= 1e-5 relative; gradients <= 1e-4 relative (inf-norm over inf-norm) on the reference's fixtures, whose sorted
errors are separated (manifest ``order_margin`` > 1) so that the gradient is independent of the soft-max implementation.
On larger inputs near-tied errors are unavoidable (tens of thousands of fp32 values in [0,1] per segment) and the
piecewise-constant gradient of the two tied pixels depends on their order — in the reference too (torch.sort's order among
ties is unspecified). There the gradient is checked (a) element-wise against the exact (float64 Jaccard) oracle with at most
1e-4 of the elements (or four flipped pairs) outside the 1e-4 gate and none outside 1e-2, and (b) through an order-independent property: the
directional derivative of the oracle's loss along sign(grad).
"""
import warnings

import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.helpers import rel_err, synth_labels, synth_logits

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


def _run(B, x, y, kw, avg_factor=None, ignore=255, grad_out=None):
    x = x.detach().clone().requires_grad_(True)
    loss = B.LovaszLoss(**kw)(x, y, weight=None, avg_factor=avg_factor, ignore_index=ignore)
    if loss.dim():
        (loss * grad_out.to(loss.dtype)).sum().backward()
    else:
        loss.backward()
    return loss.detach(), x.grad


def _oracle64(x, y, kw, avg_factor=None, ignore=255, grad_out=None):
    xo = x.detach().double().requires_grad_(True)
    lo = O.lovasz_loss_module(xo, y, avg_factor=avg_factor, ignore_index=ignore, acc_dtype=torch.float64, **kw)
    if lo.dim():
        (lo * grad_out.double()).sum().backward()
    else:
        lo.backward()
    return lo.detach(), xo.grad


def test_lovasz_golden(B, golden):
    data, manifest = golden
    cases = [c for c in manifest['cases'] if c['kind'] == 'lovasz']
    assert len(cases) >= 12
    for case in cases:
        name = case['name']
        x = torch.from_numpy(data[name + '/logits']).cuda()
        y = torch.from_numpy(data[name + '/labels']).cuda()
        go = torch.from_numpy(data[name + '/grad_out']).cuda() if (name + '/grad_out') in data else None
        loss, grad = _run(B, x, y, case['kw'], case.get('avg_factor'), case['ignore'], go)
        assert tuple(loss.shape) == tuple(data[name + '/loss'].shape), name
        assert rel_err(loss, data[name + '/loss']) <= LOSS_TOL, '%s loss %.3e' % (name, rel_err(loss, data[name + '/loss']))
        assert rel_err(grad, data[name + '/grad']) <= GRAD_TOL, '%s grad %.3e' % (name, rel_err(grad, data[name + '/grad']))
        # forward only (keys-only sort) gives the same value
        with torch.no_grad():
            l2 = B.LovaszLoss(**case['kw'])(x, y, avg_factor=case.get('avg_factor'), ignore_index=case['ignore'])
        assert rel_err(l2, loss) <= 1e-6, name


def _grad_gate(name, grad, ref):
    """Element-wise gate that tolerates a handful of near-tie order flips: one flipped pair of (nearly) equal errors moves
    the gradient of its two pixels — all C components of each, through the soft-max Jacobian — by a second-order amount."""
    d = (grad.double() - ref.double()).abs()
    scale = float(ref.abs().max())
    bad = int((d > GRAD_TOL * scale).sum())
    allowed = max(int(1e-4 * d.numel()), 8 * (ref.shape[1] if ref.dim() == 4 else 1))
    worst = float(d.max()) / scale
    assert bad <= allowed and worst <= 1e-2, '%s: %d gradient elements outside %.0e (allowed %d), worst %.3e' % (
        name, bad, GRAD_TOL, allowed, worst)


@pytest.mark.parametrize('shape,C,kw', [
    ((4, 19, 96, 160), 19, dict(reduction='none')),
    ((2, 150, 32, 64), 150, dict(reduction='none', class_weight=np.linspace(0.5, 1.5, 150).tolist())),
    ((3, 7, 37, 53), 7, dict(reduction='none', classes='all')),                     # odd extent: scalar kernels
    ((4, 6, 64, 64), 6, dict(per_image=True, reduction='mean', loss_weight=0.5)),
    ((3, 5, 33, 31), 5, dict(per_image=True, reduction='none', classes=[0, 2, 4])),
])
def test_lovasz_softmax_vs_exact_oracle(B, shape, C, kw):
    x = synth_logits(shape, 41, device='cuda', margin=False)
    y = synth_labels((shape[0],) + shape[2:], C, 41, ignore_index=255, block=8, device='cuda')
    go = torch.linspace(0.5, 1.5, shape[0], device='cuda') if (kw.get('per_image') and kw['reduction'] == 'none') else None
    loss, grad = _run(B, x, y, kw, grad_out=go)
    lo, go_ref = _oracle64(x, y, kw, grad_out=go)
    assert rel_err(loss, lo) <= LOSS_TOL, '%s loss %.3e' % (shape, rel_err(loss, lo))
    _grad_gate(str(shape), grad, go_ref)
    # order-independent: directional derivative of the exact loss along sign(grad) (a random direction projects the
    # gradient onto almost nothing, and the central difference then mostly measures the kinks of the piecewise-linear
    # Lovasz extension that lie within +-eps)
    d = torch.sign(go_ref).float()
    eps = 1e-4
    sel = (lambda l: (l * go.double()).sum()) if go is not None else (lambda l: l)
    with torch.no_grad():
        lp = sel(O.lovasz_loss_module((x.double() + eps * d.double()), y, ignore_index=255, acc_dtype=torch.float64, **kw))
        lm = sel(O.lovasz_loss_module((x.double() - eps * d.double()), y, ignore_index=255, acc_dtype=torch.float64, **kw))
    fd = float((lp - lm) / (2 * eps))
    an = float((grad.double() * d.double()).sum())
    assert abs(fd - an) <= 2e-3 * abs(fd), '%s directional derivative %.6e vs %.6e' % (shape, an, fd)
    # ignored pixels receive no gradient; the soft-max Jacobian sums to zero over the classes
    ign = (y == 255).unsqueeze(1).expand_as(grad)
    assert float(grad[ign].abs().max()) == 0.0
    assert float(grad.double().sum(1).abs().max()) <= 1e-5 * float(grad.abs().max())


@pytest.mark.parametrize('kw', [dict(loss_type='binary', reduction='none'),
                                dict(loss_type='binary', per_image=True, reduction='mean'),
                                dict(loss_type='binary', per_image=True, reduction='sum', loss_weight=2.0)])
def test_lovasz_hinge_vs_exact_oracle(B, kw):
    g = torch.Generator().manual_seed(3)
    x = (torch.randn((4, 1, 48, 80), generator=g) * 2).cuda()
    y = synth_labels((4, 48, 80), 2, 3, ignore_index=255, block=4, device='cuda')
    loss, grad = _run(B, x, y, kw)
    lo, gr = _oracle64(x, y, kw)
    assert rel_err(loss, lo) <= LOSS_TOL, rel_err(loss, lo)
    _grad_gate('hinge', grad, gr)
    x3 = x.squeeze(1)                                      # (N,H,W) logits, the documented binary layout (:99-100)
    l3, g3 = _run(B, x3, y, kw)
    assert rel_err(l3, loss) <= 1e-6 and rel_err(g3, grad.squeeze(1)) <= 1e-6


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
def test_lovasz_half_logits(B, dtype):
    shape, C = (2, 21, 64, 96), 21
    x = synth_logits(shape, 17, dtype=dtype, device='cuda')
    y = synth_labels((2, 64, 96), C, 17, ignore_index=255, block=8, device='cuda')
    loss, grad = _run(B, x, y, dict(reduction='none'))
    assert loss.dtype == dtype and grad.dtype == dtype
    lo, gr = _oracle64(x.float(), y, dict(reduction='none'))
    assert rel_err(loss, lo) <= HALF_TOL
    d = (grad.double() - gr).abs()
    assert float((d > 2 * HALF_TOL * float(gr.abs().max())).double().mean()) <= 1e-4


@pytest.mark.parametrize('label_dtype', [torch.uint8, torch.int32, torch.float32])
def test_lovasz_label_dtypes_and_edge_cases(B, label_dtype):
    shape, C = (2, 5, 24, 40), 5
    x = synth_logits(shape, 23, device='cuda', margin=False)
    y = synth_labels((2, 24, 40), C, 23, ignore_index=255, block=4, device='cuda')
    l0, g0 = _run(B, x, y, dict(reduction='none'))
    l1, g1 = _run(B, x, y.to(label_dtype), dict(reduction='none'))
    assert torch.equal(g0, g1) and rel_err(l1, l0) <= 1e-6
    # (N,1,H,W) labels, as the decode head holds them before squeeze(1)
    l2, _ = _run(B, x, y.unsqueeze(1), dict(reduction='none'))
    assert rel_err(l2, l0) <= 1e-6
    # every pixel ignored: zero loss, zero gradient (the reference's stated intent, :148-150)
    ya = torch.full_like(y, 255)
    la, ga = _run(B, x, ya, dict(reduction='none'))
    assert float(la) == 0.0 and float(ga.abs().max()) == 0.0
    lb, gb = _run(B, x, ya, dict(per_image=True, reduction='mean'))
    assert float(lb) == 0.0 and float(gb.abs().max()) == 0.0
    # one image fully ignored under per_image: it contributes a zero to the mean
    yb = y.clone()
    yb[1] = 255
    lc, _ = _run(B, x, yb, dict(per_image=True, reduction='mean'))
    ld, _ = _run(B, x[:1], y[:1], dict(per_image=True, reduction='mean'))
    assert rel_err(lc * 2, ld) <= 1e-6
    # ignore_index=None: every pixel counts
    yn = y.clone()
    yn[yn == 255] = 0
    ln, gn = _run(B, x, yn, dict(reduction='none'), ignore=None)
    lo, go = _oracle64(x, yn, dict(reduction='none'), ignore=None)
    assert rel_err(ln, lo) <= LOSS_TOL
    _grad_gate('no-ignore', gn, go)


def test_lovasz_constructor_and_errors(B):
    with pytest.raises(AssertionError):
        B.LovaszLoss(loss_type='softmax')
    with pytest.raises(AssertionError):
        B.LovaszLoss(reduction='mean')                     # per_image=False needs reduction='none' (:271-273)
    with pytest.raises(AssertionError):
        B.LovaszLoss(classes='some', reduction='none')
    with pytest.raises(ValueError):
        B.LovaszLoss(classes=[1, 1, 2], reduction='none')
    m = B.LovaszLoss(per_image=True, reduction='sum')
    assert m.loss_name == 'loss_lovasz' and len(m.state_dict()) == 0
    x = torch.randn(2, 3, 8, 8, device='cuda')
    y = torch.randint(0, 3, (2, 8, 8), device='cuda')
    with pytest.raises(ValueError):
        m(x, y, avg_factor=2.0)                            # models/losses/utils.py:78-79
    with pytest.raises(AssertionError):
        m(x, y, reduction_override='avg')
    with pytest.raises(RuntimeError):
        B.LovaszLoss(reduction='none')(x.cpu(), y.cpu())   # no CPU path
    reg = B.registry.build_loss(dict(type='LovaszLoss', reduction='none', loss_weight=0.3))
    assert isinstance(reg, B.LovaszLoss) and reg.loss_weight == 0.3


def test_lovasz_cityscapes_shape_properties(B):
    """BASELINE config-2 label resolution (8 x 19 x 512 x 1024, 4 M pixels per class segment): size-independent properties."""
    N, C, H, W = 8, 19, 512, 1024
    x = synth_logits((N, C, H, W), 2, device='cuda', margin=False)
    y = synth_labels((N, H, W), C, 2, ignore_index=255, device='cuda')
    xg = x.clone().requires_grad_(True)
    loss = B.LovaszLoss(reduction='none')(xg, y, ignore_index=255)
    loss.backward()
    g = xg.grad
    assert 0.0 < float(loss) < 1.0 and bool(torch.isfinite(g).all())
    assert float(g[(y == 255).unsqueeze(1).expand_as(g)].abs().max()) == 0.0
    assert float(g.double().sum(1).abs().max()) <= 1e-5 * float(g.abs().max())
    # directional derivative against the loss itself (forward-only path, keys-only sort)
    # (along sign(grad): the fp32 loss values must differ by much more than their rounding)
    d = torch.sign(g)
    eps = 1e-3
    with torch.no_grad():
        lp = float(B.LovaszLoss(reduction='none')(x + eps * d, y, ignore_index=255).double())
        lm = float(B.LovaszLoss(reduction='none')(x - eps * d, y, ignore_index=255).double())
    fd = (lp - lm) / (2 * eps)
    an = float((g.double() * d.double()).sum())
    assert abs(fd - an) <= 1e-2 * abs(fd), (fd, an)
    # a permutation of the batch leaves the batch-level loss unchanged (sum over a multiset of pixels)
    perm = torch.tensor([3, 1, 7, 0, 2, 6, 5, 4], device='cuda')
    with torch.no_grad():
        lperm = B.LovaszLoss(reduction='none')(x[perm].contiguous(), y[perm].contiguous(), ignore_index=255)
    assert rel_err(lperm, loss) <= 1e-6
    # exact oracle on one class-sized problem would take minutes on the host; the per-image loss of one image is checked
    with torch.no_grad():
        l1 = B.LovaszLoss(per_image=True, reduction='none')(x[:1], y[:1], ignore_index=255)
    lo = O.lovasz_loss_module(x[:1].double(), y[:1], per_image=True, reduction='none', ignore_index=255, acc_dtype=torch.float64)
    assert rel_err(l1, lo) <= LOSS_TOL


def test_lovasz_in_decode_head_call(B):
    """decode_head.py:283-293: CE + Lovasz on low-resolution logits through the fused call site (the resize is materialised
    once for the Lovasz module; CE keeps its resize-fused single pass)."""
    x = synth_logits((2, 7, 16, 24), 9, device='cuda', margin=False).requires_grad_(True)
    y = synth_labels((2, 64, 96), 7, 9, ignore_index=255, block=8, device='cuda').unsqueeze(1)
    out = B.fused_resize_losses(x, y, [B.CrossEntropyLoss(), B.LovaszLoss(reduction='none', loss_weight=0.5)], ignore_index=255)
    assert list(out.keys()) == ['loss_ce', 'loss_lovasz', 'acc_seg']
    (out['loss_ce'] + out['loss_lovasz']).backward()
    xo = x.detach().double().requires_grad_(True)
    full = O.resize(xo, size=(64, 96), mode='bilinear', align_corners=False)
    ref_ce = O.cross_entropy_loss_module(full, y.squeeze(1), ignore_index=255)
    ref_lv = O.lovasz_loss_module(full, y.squeeze(1), reduction='none', loss_weight=0.5, ignore_index=255, acc_dtype=torch.float64)
    (ref_ce + ref_lv).backward()
    assert rel_err(out['loss_ce'], ref_ce) <= LOSS_TOL and rel_err(out['loss_lovasz'], ref_lv) <= LOSS_TOL
    _grad_gate('head', x.grad, xo.grad)


def test_lovasz_cuda_graph(B):
    x = synth_logits((2, 6, 32, 48), 5, device='cuda', margin=False)
    y = synth_labels((2, 32, 48), 6, 5, ignore_index=255, block=4, device='cuda')
    mod = B.LovaszLoss(reduction='none')
    xs = x.clone().requires_grad_(True)
    l_eager = mod(xs, y, ignore_index=255)
    l_eager.backward()
    g_eager = xs.grad.clone()
    xg = x.clone().requires_grad_(True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            xg.grad = None
            mod(xg, y, ignore_index=255).backward()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    xg.grad = None
    with torch.cuda.graph(graph):
        l_cap = mod(xg, y, ignore_index=255)
        l_cap.backward()
    graph.replay()
    torch.cuda.synchronize()
    assert rel_err(l_cap, l_eager) <= 1e-6 and rel_err(xg.grad, g_eager) <= 1e-6


def test_lovasz_cabi_buffers_have_no_out_of_bounds_writes(B):
    """compute-sanitizer is not available on the GPU pool: every buffer the Lovasz entries write (compact labels, G, sort
    workspace, segment statistics, outputs, coefficients, gradient) sits between canary words; vector and scalar kernels,
    partial scan tiles, per-image segments and the binary variant."""
    import ctypes as C
    from image_segmentation_lab_b200 import _lib
    lib = B.load_library()
    dev = torch.device('cuda', 0)
    stream = _lib.stream_ptr(dev)
    GUARD = 4096

    def guarded(nbytes):
        n = (nbytes + 255) // 256 * 256
        buf = torch.full((n + 2 * GUARD,), 0x5A, dtype=torch.uint8, device=dev)
        return buf, buf[GUARD:GUARD + n], n

    def intact(buf, n, what):
        assert bool((buf[:GUARD] == 0x5A).all()) and bool((buf[GUARD + n:] == 0x5A).all()), what + ': guard overwritten'

    # last field: classes the workspace is sized for (fewer than C: the classes go through in several batches of launches)
    cases = [((2, 5, 24, 40), False, False, 5), ((3, 7, 37, 53), False, False, 7), ((1, 3, 50, 41), False, False, 3),
             ((3, 4, 45, 46), False, True, 4), ((2, 1, 33, 31), True, False, 1), ((3, 1, 64, 48), True, True, 1),
             ((2, 5, 67, 93), False, False, 2), ((3, 4, 45, 46), False, True, 1)]
    for shape, binary, per_image, ws_classes in cases:
        N, Cc, H, W = shape
        HW = H * W
        x = synth_logits(shape, 13, device='cuda', margin=False)
        y = synth_labels((N, H, W), 2 if binary else Cc, 13, ignore_index=255, block=5, device='cuda')
        n_groups = N if per_image else 1
        n_seg = 1 if binary else Cc
        lse = torch.logsumexp(x.double(), 1).float().reshape(N, HW).contiguous() if not binary else None
        ws_bytes = int(lib.b200seg_lovasz_workspace_bytes(N, ws_classes, HW, int(per_image), 1))
        bufs = {}
        for name, nbytes in (('lab16', N * HW * 2), ('G', N * n_seg * HW * 4), ('ws', ws_bytes), ('seg', n_groups * n_seg * 16),
                             ('out', max(n_groups, 1) * 4), ('coef', n_groups * n_seg * 4), ('grad', N * Cc * HW * 4)):
            bufs[name] = guarded(nbytes)
        d = _lib.LovaszDesc()
        d.logits = x.data_ptr(); d.labels = y.data_ptr(); d.lse = lse.data_ptr() if lse is not None else None
        d.logit_dtype = _lib.F32; d.label_dtype = _lib.L_I64
        d.N, d.C, d.HW = N, Cc, HW
        d.ignore_index = 255; d.has_ignore = 1
        d.binary = int(binary); d.per_image = int(per_image); d.only_present = 1
        d.reduction = _lib.RED_NONE if per_image else _lib.RED_MEAN
        d.loss_weight = 1.0
        d.lab16 = bufs['lab16'][1].data_ptr(); d.G = bufs['G'][1].data_ptr()
        d.workspace = bufs['ws'][1].data_ptr(); d.workspace_bytes = ws_bytes
        assert d.workspace % 256 == 0
        d.seg_stats = bufs['seg'][1].data_ptr(); d.out = bufs['out'][1].data_ptr(); d.coef = bufs['coef'][1].data_ptr()
        _lib.check(lib.b200seg_lovasz_fwd(C.byref(d), stream))
        b = _lib.LovaszBwdDesc()
        b.logits = x.data_ptr(); b.lse = d.lse; b.lab16 = d.lab16; b.G = d.G; b.coef = d.coef
        b.grad_logits = bufs['grad'][1].data_ptr()
        b.logit_dtype = _lib.F32; b.N, b.C, b.HW = N, Cc, HW
        b.binary = int(binary); b.per_image = int(per_image); b.grad_per_group = 0
        _lib.check(lib.b200seg_lovasz_bwd(C.byref(b), stream))
        torch.cuda.synchronize()
        for name, (buf, _, n) in bufs.items():
            intact(buf, n, '%s %s' % (shape, name))
        # and the values are the oracle's
        kw = dict(loss_type='binary' if binary else 'multi_class', per_image=per_image, reduction='none')
        xo = x.double().requires_grad_(True)
        lo = O.lovasz_loss_module(xo, y, ignore_index=255, acc_dtype=torch.float64, **kw)
        lo.sum().backward()
        got = bufs['out'][1][:max(n_groups, 1) * 4].view(torch.float32)[:n_groups if per_image else 1]
        assert rel_err(got.reshape(lo.shape), lo) <= LOSS_TOL, shape
        grad = bufs['grad'][1][:N * Cc * HW * 4].view(torch.float32).reshape(shape)
        _grad_gate(str(shape), grad, xo.grad)
    # a too-small workspace is refused before any launch
    d.workspace_bytes = 1024
    assert lib.b200seg_lovasz_fwd(C.byref(d), stream) != 0 and 'workspace' in _lib.last_error()


@pytest.mark.parametrize('shape,kw', [
    ((3, 5, 157, 211), dict(reduction='none')),                              # 24 full sort tiles + a partial one per class
    ((5, 3, 91, 123), dict(per_image=True, reduction='mean')),               # 15 segments of 2.7 tiles
    ((2, 2, 300, 301), dict(reduction='none', classes='all')),
])
def test_lovasz_multi_tile_segments(B, shape, kw):
    """The hand-written radix sort across tile boundaries: look-back chains of tens of tiles, a partial last tile, several
    segments per launch; loss and gradient against the exact-Jaccard fp64 oracle."""
    x = synth_logits(shape, 31, device='cuda', margin=False)
    y = synth_labels(shape[:1] + shape[2:], shape[1], 31, ignore_index=255, block=7, device='cuda')
    l, g = _run(B, x, y, kw)
    xo = x.double().requires_grad_(True)
    lo = O.lovasz_loss_module(xo, y, ignore_index=255, acc_dtype=torch.float64, **kw)
    lo.sum().backward()
    assert rel_err(l.reshape(lo.shape), lo) <= LOSS_TOL
    _grad_gate(str(shape), g, xo.grad)


def test_lovasz_sort_ties_and_determinism(B):
    """Logits quantised to a few values: long runs of equal sort keys (the stable sort keeps pixel order inside them; the
    loss does not depend on that order, lovasz_loss.py:26-39 telescopes) — and two runs are bit-identical."""
    shape = (2, 4, 96, 128)
    x = (synth_logits(shape, 5, device='cuda', margin=False) * 2).round() / 2
    y = synth_labels(shape[:1] + shape[2:], shape[1], 5, ignore_index=255, block=9, device='cuda')
    l1, g1 = _run(B, x, y, dict(reduction='none'))
    l2, g2 = _run(B, x, y, dict(reduction='none'))
    assert torch.equal(g1, g2)
    assert abs(float(l1) - float(l2)) <= 1e-6 * abs(float(l1))          # the class sums are fp64 atomics of fp32 tile sums
    lo = O.lovasz_loss_module(x.double(), y, ignore_index=255, acc_dtype=torch.float64, reduction='none')
    assert rel_err(l1.reshape(lo.shape), lo) <= LOSS_TOL
    z = torch.zeros(shape, device='cuda')                                   # every key of a class equal
    lz, _ = _run(B, z, y, dict(reduction='none'))
    lzo = O.lovasz_loss_module(z.double(), y, ignore_index=255, acc_dtype=torch.float64, reduction='none')
    assert rel_err(lz.reshape(lzo.shape), lzo) <= LOSS_TOL
