"""Build libb2ingest.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python image-classification-system_b200/build.py [--force] [--verbose]

The shared library is written next to this file so it travels to the GPU box with the
repository snapshot.  Sources are compiled one object per .cu (in parallel) and linked with
the static CUDA runtime, so the .so depends only on libcuda (driver) at run time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb2ingest.so")
SOURCES = ["api.cu", "sha256.cu", "dedupe.cu", "resize.cu", "tally.cu", "host.cu", "ring.cu", "comm.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "b2ingest.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr", "--extended-lambda",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and not _stale(LIB, srcs + HEADERS + [os.path.abspath(__file__)]):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + HEADERS + [os.path.abspath(__file__)]):
            cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = os.path.join(BUILD, os.path.basename(src)[:-3] + ".ptxas.log")
            with open(log, "w") as f:
                f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src} (see {log})")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-cudart", "static", "-lcuda", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libb2ingest.so failed")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
