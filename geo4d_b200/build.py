"""In-tree build of libgeo4d_b200.so (sm_100a only) with plain nvcc.

    python -m geo4d_b200.build [--force] [--verbose]

The shared library is the C-ABI product boundary (include/geo4d_b200.h); it
links only the CUDA runtime (static) and resolves the one driver symbol it
needs (cuTensorMapEncodeTiled) at run time, so it loads on a CPU-only box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(REPO, "include")
BUILD_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "libgeo4d_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                "-I", INCLUDE, "-I", CSRC]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path, extra=()):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    for hdr in sorted(os.listdir(CSRC)):
        if hdr.endswith((".cuh", ".h")):
            with open(os.path.join(CSRC, hdr), "rb") as f:
                h.update(f.read())
    with open(os.path.join(INCLUDE, "geo4d_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(COMMON_FLAGS + ARCH_FLAGS + list(extra)).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    objs, jobs = [], []
    for src in sources():
        sp = os.path.join(CSRC, src)
        obj = os.path.join(BUILD_DIR, src[:-3] + ".o")
        stamp = obj + ".sha"
        dg = _digest(sp)
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.exists(stamp)
                and open(stamp).read() == dg):
            continue
        cmd = [NVCC, *ARCH_FLAGS, *COMMON_FLAGS, "-c", sp, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        jobs.append((cmd, stamp, dg))

    def run(job):
        cmd, stamp, dg = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose or r.stderr.strip():
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(dg)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or force or not os.path.exists(LIB_PATH):
        cmd = [NVCC, *ARCH_FLAGS, "-shared", "-o", LIB_PATH, *objs, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
