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


def build_variant(name, defines):
    """Experimental variant: libnttb200_<name>.so compiled with extra -D flags (selected at run time by NTTB200_LIB)."""
    objdir = os.path.join(OBJ, name)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in _sources():
        obj = os.path.join(objdir, src[:-3] + ".o")
        subprocess.check_call([NVCC] + FLAGS + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj])
        objs.append(obj)
    out = os.path.join(HERE, "nttb200", "libnttb200_%s.so" % name)
    subprocess.check_call([NVCC, "-shared", "-o", out] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"])
    return out


def build(force=False, jobs=8):
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), _sources()))
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(o) for o in objs):
        subprocess.check_call([NVCC, "-shared", "-o", OUT] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"])
    return OUT


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv))
