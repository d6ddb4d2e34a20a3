"""In-tree build of libmtt_b200.so (CUDA kernels + C-ABI) for sm_100a.

The shared library is built next to this file so that it travels to the GPU box with the
repository snapshot.  Only nvcc and g++ are needed; nothing is downloaded.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD = os.path.join(PKG_DIR, "csrc", "_build")
LIB = os.path.join(PKG_DIR, "libmtt_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2", "--expt-relaxed-constexpr",
]
CXX_FLAGS = ["-O2", "-g", "-fPIC", "-std=c++17", "-ffp-contract=off"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _sources():
    cu, cpp = [], []
    for dirpath, _, files in os.walk(CSRC):
        if "_build" in dirpath:
            continue
        for f in sorted(files):
            if f.endswith(".cu"):
                cu.append(os.path.join(dirpath, f))
            elif f.endswith(".cpp"):
                cpp.append(os.path.join(dirpath, f))
    return cu, cpp


def _headers_mtime() -> float:
    newest = 0.0
    for dirpath, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cuh", ".h", ".hpp")):
                newest = max(newest, os.path.getmtime(os.path.join(dirpath, f)))
    newest = max(newest, os.path.getmtime(os.path.join(ROOT, "include", "mtt_b200.h")))
    return newest


def _compile(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA / C++ source under csrc/ and link libmtt_b200.so. Returns its path."""
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    cu, cpp = _sources()
    hdr_m = _headers_mtime()
    jobs, objs = [], []
    for src in cu + cpp:
        obj = os.path.join(BUILD, os.path.relpath(src, CSRC).replace(os.sep, "_") + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(
            os.path.getmtime(src), hdr_m)
        if not stale:
            continue
        if src.endswith(".cu"):
            jobs.append([nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-c", src,
                         "-o", obj])
        else:
            jobs.append(["g++", *CXX_FLAGS, "-I", os.path.join(ROOT, "include"), "-c", src,
                         "-o", obj])
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(_compile, jobs):
                if verbose and out:
                    print(out, file=sys.stderr)
    if jobs or not os.path.exists(LIB):
        _compile([nvcc, "-shared", "-o", LIB, *objs, "-gencode",
                  "arch=compute_100a,code=sm_100a", "-lcudart_static", "-lpthread", "-ldl",
                  "-lrt"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
