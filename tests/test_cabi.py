"""CPU: the C-ABI library loads, exports every symbol include/b200seg.h declares, and the ctypes
structures have the layout the C compiler gives them (no compute calls without a GPU)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'b200seg.h')


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = re.findall(r'\b(b200seg_[a-z0-9_]+)\s*\(', src)
    return sorted(set(names))


def test_header_declares_entry_points():
    names = _declared_functions()
    for must in ('b200seg_loss_fwd', 'b200seg_loss_bwd', 'b200seg_loss_finalize', 'b200seg_loss_fused_fwdbwd',
                 'b200seg_confusion_labels', 'b200seg_confusion_logits', 'b200seg_resize_bilinear_fwd',
                 'b200seg_resize_bilinear_bwd', 'b200seg_last_error'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from image_segmentation_lab_b200 import _lib
    lib = _lib.load()
    assert os.path.exists(_lib.lib_path())
    for name in _declared_functions():
        assert hasattr(lib, name), 'libb200seg.so does not export %s' % name
    assert lib.b200seg_abi_version() == _lib.ABI_VERSION
    assert lib.b200seg_confusion_chunk_pixels() > 0
    assert lib.b200seg_launch_count() >= 0


def test_ctypes_bindings_cover_header():
    from image_segmentation_lab_b200 import _lib
    bound = {s[0] for s in _lib.SYMBOLS}
    for name in _declared_functions():
        assert name in bound, '%s is declared in the header but not bound in _lib.py' % name


def test_struct_layouts_match_c_compiler():
    from image_segmentation_lab_b200 import _lib
    gcc = 'gcc'
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "b200seg.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(b200seg_loss_desc), sizeof(b200seg_finalize_desc), sizeof(b200seg_loss_bwd_desc),
         sizeof(b200seg_loss_fused_desc), sizeof(b200seg_image), sizeof(b200seg_bce_desc));
  printf("%zu %zu %zu %zu\n", offsetof(b200seg_loss_desc, ignore_index), offsetof(b200seg_loss_desc, stats),
         offsetof(b200seg_loss_bwd_desc, grad_logits), offsetof(b200seg_loss_fused_desc, workspace));
  printf("%zu %zu %zu %zu %zu\n", sizeof(b200seg_lovasz_desc), sizeof(b200seg_lovasz_bwd_desc), offsetof(b200seg_lovasz_desc, avg_factor),
         offsetof(b200seg_lovasz_desc, coef), offsetof(b200seg_lovasz_bwd_desc, HW));
  return 0;
}
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, 't.c')
        exe = os.path.join(d, 't')
        open(src, 'w').write(prog)
        subprocess.run([gcc, '-I', os.path.join(ROOT, 'include'), src, '-o', exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes[:6] == [ctypes.sizeof(_lib.LossDesc), ctypes.sizeof(_lib.FinalizeDesc), ctypes.sizeof(_lib.LossBwdDesc),
                         ctypes.sizeof(_lib.LossFusedDesc), ctypes.sizeof(_lib.Image), ctypes.sizeof(_lib.BceDesc)]
    assert sizes[6:10] == [_lib.LossDesc.ignore_index.offset, _lib.LossDesc.stats.offset,
                           _lib.LossBwdDesc.grad_logits.offset, _lib.LossFusedDesc.workspace.offset]
    assert sizes[10:] == [ctypes.sizeof(_lib.LovaszDesc), ctypes.sizeof(_lib.LovaszBwdDesc), _lib.LovaszDesc.avg_factor.offset,
                          _lib.LovaszDesc.coef.offset, _lib.LovaszBwdDesc.HW.offset]


def test_validation_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on a CPU-only box."""
    from image_segmentation_lab_b200 import _lib
    lib = _lib.load()
    d = _lib.LossDesc()
    d.N, d.C, d.h, d.w, d.H, d.W = 1, 0, 4, 4, 4, 4
    assert lib.b200seg_loss_fwd(ctypes.byref(d), None) != 0
    assert 'bad shape' in _lib.last_error()
    f = _lib.FinalizeDesc()
    assert lib.b200seg_loss_finalize(ctypes.byref(f), None) != 0
    lv = _lib.LovaszDesc()
    lv.N, lv.C, lv.HW, lv.binary = 2, 3, 16, 1
    assert lib.b200seg_lovasz_fwd(ctypes.byref(lv), None) != 0
    assert 'single-channel' in _lib.last_error()
    lv.binary, lv.per_image, lv.has_avg_factor, lv.reduction = 0, 1, 1, _lib.RED_SUM
    assert lib.b200seg_lovasz_fwd(ctypes.byref(lv), None) != 0
    assert 'avg_factor can not be used' in _lib.last_error()
    assert lib.b200seg_loss_fused_workspace_bytes(8, 19, 64, 128, 512, 1024, 0) == 8 * 19 * 65 * 129 * 16
    assert lib.b200seg_loss_fused_workspace_bytes(8, 19, 64, 128, 512, 1024, 1) == 8 * 19 * 65 * 129 * 16   # any align_corners
    assert lib.b200seg_loss_fused_workspace_bytes(2, 19, 65, 129, 513, 1025, 1) == 2 * 19 * 66 * 130 * 16   # any up-sampling ratio
    assert lib.b200seg_loss_fused_workspace_bytes(8, 150, 64, 64, 512, 512, 0) == 8 * 150 * 65 * 65 * 16 + 8 * 512 * 512 * 4   # class-tiled
    assert lib.b200seg_loss_fused_workspace_bytes(8, 600, 64, 64, 512, 512, 0) == 0    # C > 512 -> resize first
    assert lib.b200seg_loss_fused_workspace_bytes(8, 19, 64, 128, 32, 1024, 0) == 0    # down-sampling -> two-pass path


def test_sass_is_sm100a():
    """The shipped library carries sm_100a SASS only (no PTX-JIT / multi-arch fallback)."""
    from image_segmentation_lab_b200 import _lib
    cuobjdump = '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([cuobjdump, '-lelf', _lib.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs
