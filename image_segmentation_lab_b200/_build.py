"""Builds libb200seg.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

The shared library is the product: there is no JIT cache and no fallback. It is built next to this
file so that it travels with the source tree (it is git-ignored, not gpurun-ignored).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_PATH = os.path.join(HERE, "libb200seg.so")
SOURCES = ["api.cu", "loss_stream.cu", "loss_rt.cu", "loss_rt_f32.cu", "loss_rt_bf16.cu", "loss_rt_f16.cu", "loss_up.cu", "loss_upcell_f32.cu", "loss_upcell_bf16.cu", "loss_upcell_f16.cu", "loss_upgen_f32.cu", "loss_upgen_bf16.cu", "loss_upgen_f16.cu",
           "loss_dice.cu", "loss_cs.cu", "loss_bulk.cu", "loss_bce.cu", "loss_lovasz.cu", "resize.cu", "confusion.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200seg.so cannot be built (set NVCC=/path/to/nvcc)")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def source_digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    files.append(os.path.join(INCLUDE, "b200seg.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    stamp = LIB_PATH + ".stamp"
    if not (os.path.exists(LIB_PATH) and os.path.exists(stamp)):
        return False
    with open(stamp) as fh:
        return fh.read().strip() == source_digest()


def build_library(force=False, verbose=False, ptxas_verbose=False):
    """Compile every .cu for sm_100a and link libb200seg.so. Returns the library path."""
    if not force and is_current():
        return LIB_PATH
    nvcc = _nvcc()
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    flags = list(NVCC_FLAGS) + (["-Xptxas", "-v"] if ptxas_verbose else [])

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc] + flags + ["-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if ptxas_verbose:
            with open(obj + ".ptxas.log", "w") as fh:
                fh.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-4000:]))
        if verbose:
            print("[b200seg] compiled", src, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stderr[-4000:])
    with open(LIB_PATH + ".stamp", "w") as fh:
        fh.write(source_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True, ptxas_verbose="--ptxas" in sys.argv))
