#!/usr/bin/env python
"""Build the native sm_100a runtime in-tree: csrc/*.cu -> distkeras_b200/lib/libdistkeras_b200.so.

nvcc cross-compiles for sm_100a without a GPU.  Objects are cached under build/obj and only
rebuilt when a source or header changed.  `python build_native.py --tests` also builds the
standalone C++/CUDA test binaries under build/.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB_DIR = os.path.join(ROOT, "distkeras_b200", "lib")
LIB = os.path.join(LIB_DIR, "libdistkeras_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "--extended-lambda", "-Xcompiler", "-fPIC",
                     "-Xcompiler", "-fvisibility=default"]

SOURCES = ["gemm_tcgen05.cu", "ps_kernels.cu", "optim_kernels.cu", "loss_kernels.cu", "nn_kernels.cu",
           "fabric.cu", "dense_fused.cu", "conv_tma.cu", "engine.cu"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]


def compile_one(src: str, verbose: bool) -> str:
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    stamp = obj + ".sha"
    dig = _digest([path] + _headers())
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj
    cmd = [nvcc()] + NVCC_FLAGS + ["-I", CSRC, "-c", path, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed for {src}")
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return obj


def build(verbose: bool = False, tests: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: compile_one(s, verbose), SOURCES))
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    if tests:
        tdir = os.path.join(CSRC, "tests")
        for t in sorted(os.listdir(tdir)):
            if not t.endswith(".cu"):
                continue
            out = os.path.join(ROOT, "build", t[:-3])
            cmd = [nvcc()] + ARCH + ["-lineinfo", "-O3", "-std=c++17", "--extended-lambda", "-I", CSRC,
                                     os.path.join(tdir, t), "-o", out,
                                     "-L", LIB_DIR, "-ldistkeras_b200", "-Xlinker", "-rpath", "-Xlinker",
                                     "$ORIGIN/../distkeras_b200/lib"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"test build failed: {t}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("--tests", action="store_true")
    a = ap.parse_args()
    print(build(a.verbose, a.tests))
