"""Ahead-of-time build of ``libseedvc_b200.so`` (nvcc, sm_100a only).

The reference JIT-builds its one CUDA op at import time for sm_70/sm_80
(modules/bigvgan/alias_free_activation/cuda/load.py:17-65); here the library is
built in-tree, explicitly, for ``compute_100a`` so it travels with the repo.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libseedvc_b200.so")
SOURCES = ["gemm.cu", "attention.cu", "elementwise.cu", "snake.cu", "hift.cu", "graph.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return "nvcc"


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_trace() -> str:
    """Debug variant with -DSVC_TRACE (clock64 event stamps); separate .so, never used by the product."""
    out = os.path.join(PKG, "libseedvc_b200_trace.so")
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DSVC_TRACE", "-shared", "-o", out] + \
        [os.path.join(CSRC, s) for s in SOURCES] + ["-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    return out


def build_variant(tag: str, defines, sources=("attention.cu",)) -> str:
    """Experiment build: recompile ``sources`` with extra -D flags, link with the regular objects of
    the other files into libseedvc_b200_<tag>.so (select it with SEEDVC_B200_LIB). Never shipped."""
    build()
    out = os.path.join(PKG, f"libseedvc_b200_{tag}.so")
    objs = []
    for s in SOURCES:
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        if s in sources:
            obj = os.path.join(CSRC, s.replace(".cu", f".{tag}.o"))
            cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + \
                [f"-D{d}" for d in defines] + ["-c", os.path.join(CSRC, s), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(r.stderr)
        objs.append(obj)
    r = subprocess.run([_nvcc(), "-shared", "-o", out] + objs +
                       ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "seedvc_b200.h"))
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append([_nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        logs = list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([_nvcc(), "-shared", "-o", LIB] + objs +
            ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    if jobs:
        with open(os.path.join(CSRC, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:      # --variant <tag> [--sources a.cu,b.cu] DEFINE ...
        i = sys.argv.index("--variant")
        rest = sys.argv[i + 2:]
        srcs = ("attention.cu",)
        if rest and rest[0] == "--sources":
            srcs, rest = tuple(rest[1].split(",")), rest[2:]
        print(build_variant(sys.argv[i + 1], rest, srcs))
    elif "--trace" in sys.argv:
        print(build_trace())
    else:
        print(build(force="--force" in sys.argv, verbose=True))
