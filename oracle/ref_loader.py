"""Loads the reference's own hot-path files BY FILE PATH (test infrastructure).

The reference cannot be imported as a package (core/__init__.py:7 imports a missing core.sampler; mmcv,
albumentations and prettytable are absent), but the files on the path load verbatim once four trivial stand-ins are
seeded in sys.modules. Two places to load them from:
  * /root/reference (this container): the sources where they lie — used to generate tests/golden/ and by
    tests/test_reference_live.py;
  * oracle/_ref/ (byte code compiled from those sources by oracle/build_ref.py; git-ignored, travels to the GPU box):
    the same unmodified code where /root/reference does not exist — the CPU arm of bench.py.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

from . import build_ref

REF_ROOT = os.environ.get('B200SEG_REFERENCE_ROOT', '/root/reference')


def source_available():
    return os.path.isfile(os.path.join(REF_ROOT, 'models', 'losses', 'cross_entropy_loss.py'))


def available():
    return source_available() or build_ref.usable()


def origin():
    """'source' (/root/reference), 'bytecode' (oracle/_ref) or None."""
    return 'source' if source_available() else ('bytecode' if build_ref.usable() else None)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _load(name, rel):
    src = os.path.join(REF_ROOT, rel)
    if os.path.isfile(src):
        spec = importlib.util.spec_from_file_location(name, src)
    else:
        pyc = build_ref.compiled_path(rel)
        spec = importlib.util.spec_from_file_location(name, pyc, loader=importlib.machinery.SourcelessFileLoader(name, pyc))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_CACHE = None


def load():
    """Returns a namespace with the reference's resize, CrossEntropyLoss, DiceLoss, accuracy, SegEvaluator, ..."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    if not available():
        raise RuntimeError('reference tree not found at %s and no byte code under oracle/_ref' % REF_ROOT)
    saved = {k: sys.modules.get(k) for k in ('mmcv', 'prettytable', 'models', 'models.builder', 'models.losses',
                                              'models.losses.utils', 'core', 'core.fileio')}

    class _Loss:
        @staticmethod
        def register(name=None):
            return lambda cls: cls

    class _PrettyTable:
        def __init__(self):
            self.cols = []

        def add_column(self, key, val):
            self.cols.append((key, list(val)))

        def get_string(self):
            return '\n'.join('%s: %s' % (k, v) for k, v in self.cols)

    sys.modules['mmcv'] = _stub('mmcv', is_list_of=lambda seq, t: isinstance(seq, list) and all(isinstance(x, t) for x in seq))
    sys.modules['prettytable'] = _stub('prettytable', PrettyTable=_PrettyTable)
    models = _stub('models')
    models.__path__ = []
    sys.modules['models'] = models
    sys.modules['models.builder'] = _stub('models.builder', LOSS=_Loss)
    losses_pkg = _stub('models.losses')
    losses_pkg.__path__ = []
    sys.modules['models.losses'] = losses_pkg
    core = _stub('core')
    core.__path__ = []
    sys.modules['core'] = core
    sys.modules['core.fileio'] = _stub('core.fileio', mkdir_or_exist=lambda d, mode=0o777: os.makedirs(d, exist_ok=True))
    try:
        ops = _load('_ref_utils_ops', 'utils/ops.py')
        lutils = _load('models.losses.utils', 'models/losses/utils.py')
        ce = _load('_ref_cross_entropy_loss', 'models/losses/cross_entropy_loss.py')
        dice = _load('_ref_dice_loss', 'models/losses/dice_loss.py')
        tv = _load('_ref_tversky_loss', 'models/losses/tversky_loss.py')
        lov = _load('_ref_lovasz_loss', 'models/losses/lovasz_loss.py')
        acc = _load('_ref_accuracy', 'models/losses/accuracy.py')
        met = _load('_ref_metrics', 'core/evaluation/metrics.py')
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    ns = types.SimpleNamespace(
        resize=ops.resize, Upsample=ops.Upsample, add_prefix=ops.add_prefix,
        reduce_loss=lutils.reduce_loss, weight_reduce_loss=lutils.weight_reduce_loss, weighted_loss=lutils.weighted_loss,
        cross_entropy=ce.cross_entropy, binary_cross_entropy=ce.binary_cross_entropy, CrossEntropyLoss=ce.CrossEntropyLoss,
        DiceLoss=dice.DiceLoss, TverskyLoss=tv.TverskyLoss, LovaszLoss=lov.LovaszLoss, lovasz_grad=lov.lovasz_grad, accuracy=acc.accuracy, Accuracy=acc.Accuracy, SegEvaluator=met.SegEvaluator)
    _CACHE = ns
    return ns


def intersect_and_union_cpu(ref, pred_labels, labels_gt, num_classes, ignore_index):
    """The reference hard-codes ``.cuda()`` (core/evaluation/metrics.py:246); run it on CPU by patching
    Tensor.cuda to the identity for the duration of the call. Inputs are copied (the reference mutates them)."""
    import torch
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        return ref.SegEvaluator.intersect_and_union(list(pred_labels), list(labels_gt), num_classes, ignore_index)
    finally:
        torch.Tensor.cuda = orig
