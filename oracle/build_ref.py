"""Recipe for oracle/_ref/: the reference's OWN hot-path files compiled to CPython byte code (test infrastructure).

The reference is Python, so "building" it means `py_compile`: every file on the path is compiled from
the source WHERE IT LIES under /root/reference into oracle/_ref/<same relative path, .refbc>. No reference source is copied
into this repository: oracle/_ref/ holds binaries only and is git-ignored (not gpurun-ignored, so it travels to the GPU
box like the built .so files, whose interpreter is the same image's CPython). oracle/ref_loader.py loads these files
when /root/reference itself is absent; `bench.py --impl reference` and the `cpu_baseline` leg then time the
UNMODIFIED reference (`cpu_baseline.kind = "reference"`) instead of the oracle port.

    python -m oracle.build_ref            # also run by __graft_entry__.build() when /root/reference exists
"""
import importlib.util
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
REF_ROOT = os.environ.get('B200SEG_REFERENCE_ROOT', '/root/reference')
# the files ref_loader.load() executes (the functions rows a1-a9 and f1/f4 of SURVEY.md 8 name)
FILES = ('utils/ops.py', 'models/losses/utils.py', 'models/losses/cross_entropy_loss.py', 'models/losses/dice_loss.py',
         'models/losses/tversky_loss.py', 'models/losses/lovasz_loss.py', 'models/losses/accuracy.py',
         'core/evaluation/metrics.py')


def compiled_path(rel):
    return os.path.join(OUT, rel[:-3] + '.refbc')   # not '.pyc': snapshot tools drop those


def build(verbose=False):
    """Returns the number of files compiled (0 when the reference tree is not on this machine)."""
    if not os.path.isfile(os.path.join(REF_ROOT, FILES[0])):
        return 0
    for rel in FILES:
        dst = compiled_path(rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(os.path.join(REF_ROOT, rel), cfile=dst, dfile='reference:' + rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(OUT, 'MAGIC'), 'w') as fh:      # byte code is tied to the interpreter version
        fh.write(importlib.util.MAGIC_NUMBER.hex())
    if verbose:
        print('[oracle/_ref] %d reference files compiled to byte code' % len(FILES), file=sys.stderr)
    return len(FILES)


def usable():
    """True when oracle/_ref holds byte code this interpreter can execute."""
    try:
        with open(os.path.join(OUT, 'MAGIC')) as fh:
            ok = fh.read().strip() == importlib.util.MAGIC_NUMBER.hex()
    except OSError:
        return False
    return ok and all(os.path.isfile(compiled_path(rel)) for rel in FILES)


if __name__ == '__main__':
    n = build(verbose=True)
    if n == 0:
        print('reference tree not found at %s: nothing built' % REF_ROOT, file=sys.stderr)
