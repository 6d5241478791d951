#!/usr/bin/env python3
"""Build libqtet_b200.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the library has a plain C ABI and is loaded with ctypes).

    python quantum-targeted-energy-transfer_b200/build_native.py [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# debug variants: QTET_NVCC_DEFS="-DQTET_PHASE_TIMING" QTET_LIB_SUFFIX=_dbg python build_native.py
SUFFIX = os.environ.get("QTET_LIB_SUFFIX", "")
EXTRA = os.environ.get("QTET_NVCC_DEFS", "").split()
OBJ = os.path.join(HERE, "build" + SUFFIX)
LIB = os.path.join(HERE, "lib", f"libqtet_b200{SUFFIX}.so")
SOURCES = ["qtet_capi.cu", "qtet_generic.cu", "qtet_split.cu", "qtet_direct.cu", "qtet_big.cu", "qtet_bigreg.cu", "qtet_hhfrag.cu", "qtet_adam.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = nvcc_path()
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "qtet.h"))
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *ARCH, *FLAGS, *EXTRA, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            jobs.append(cmd)
    if jobs:                                   # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            for r in pool.map(lambda c: subprocess.run(c, check=False), jobs):
                if r.returncode:
                    raise subprocess.CalledProcessError(r.returncode, r.args)
    if force or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
