"""Build libafb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m acquisition_focus_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repo snapshot.  No torch headers are involved: the ABI is plain C (include/afb200.h).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libafb200.so")
SOURCES = ["afb_slice.cu", "afb_views.cu", "afb_embed.cu", "afb_misc.cu", "afb_onehot.cu", "afb_minmask.cu", "afb_aux.cu", "afb_peer.cu", "afb_host.cpp"]
HEADERS = [os.path.join(CSRC, "afb_device.cuh"), os.path.join(CSRC, "afb_sampler.cuh"), os.path.join(ROOT, "include", "afb200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in _sources() + HEADERS:
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "libafb200.stamp")
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == fp:
        return LIB
    nvcc = _nvcc()
    objs = []
    for src in _sources():
        obj = os.path.join(LIBDIR, os.path.basename(src) + ".o")
        cmd = [nvcc] + ARCH + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        objs.append(obj)
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
    subprocess.run(cmd, check=True)
    with open(stamp, "w") as fh:
        fh.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
