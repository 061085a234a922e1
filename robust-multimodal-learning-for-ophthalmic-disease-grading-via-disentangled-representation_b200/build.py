"""Build the C-ABI library ``lib/libedrl_b200.so`` from ``csrc/*.cu`` with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting
``.so`` is git-ignored but travels to the GPU box with the working tree.  The library links
the CUDA runtime statically and reaches ``cuTensorMapEncodeTiled`` through
``cudaGetDriverEntryPoint`` -- no torch, no libcuda at link time.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libedrl_b200.so")
STAMP = os.path.join(LIBDIR, "libedrl_b200.stamp")
SOURCES = ["common.cu", "mmd.cu", "eprl.cu", "dilr.cu", "views.cu", "head.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "mmd_prep.cuh", "mmd_fwd.cuh", "mmd_bwd.cuh", "mmd_sweep.cuh", "topk_sift.cuh",
           os.path.join("..", "..", "include", "edrl_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]   # fast-math stays OFF: parity needs IEEE div/sqrt


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libedrl_b200.so")


def _digest() -> str:
    h = hashlib.sha256()
    for rel in SOURCES + HEADERS:
        with open(os.path.join(CSRC, rel), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if stale) and return the path of the shared library."""
    os.makedirs(LIBDIR, exist_ok=True)
    dig = _digest()
    if not force and os.path.isfile(LIB) and os.path.isfile(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == dig:
                return LIB
    nvcc = _nvcc()
    objs = []
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + (out or ""))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(STAMP, "w") as f:
        f.write(dig + "\n")
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
