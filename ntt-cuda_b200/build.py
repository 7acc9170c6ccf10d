"""Builds libnttb200.so (sm_100a) in-tree with nvcc.  `python ntt-cuda_b200/build.py [-j N] [--force]`."""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "nttb200", "libnttb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden",
         "-I", os.path.join(HERE, "..", "include")]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, force):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    srcp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(srcp), _deps_mtime()):
        return obj
    subprocess.check_call([NVCC] + FLAGS + ["-c", srcp, "-o", obj])
    return obj


def build(force=False, jobs=8):
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), _sources()))
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(o) for o in objs):
        subprocess.check_call([NVCC, "-shared", "-o", OUT] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
