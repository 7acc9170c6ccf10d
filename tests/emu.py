"""ctypes wrapper of the CPU emulator build of the CUDA kernel sources (csrc/emu).  Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "ntt-cuda_b200", "csrc", "emu")
LIB = os.path.join(EMU_DIR, "libnttb200_emu.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(EMU_DIR, f) for f in os.listdir(EMU_DIR) if f.endswith((".cpp", ".h"))]
        srcs += [os.path.join(EMU_DIR, "..", f) for f in os.listdir(os.path.join(EMU_DIR, "..")) if f.endswith(".cuh")]
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["g++", "-std=c++20", "-O2", "-DNTTB200_EMU", "-fPIC", "-shared", "-pthread", "-o", LIB,
                                   os.path.join(EMU_DIR, "emu.cpp")])
        _lib = C.CDLL(LIB)
        assert _lib.emu_sizeof_limbconst() == 80
    return _lib


def shoup(w, q):
    return np.array([(int(x) << 64) // q for x in w], dtype=np.uint64)


def limb_consts(qs, n, psiinv_tables):
    """LimbConst[limbs] as laid out in csrc/modarith.cuh (64 bytes each)."""
    out = np.zeros((len(qs), 10), dtype=np.uint64)
    for l, q in enumerate(qs):
        q = int(q)
        qbit = q.bit_length()
        mu = (1 << (2 * qbit)) // q
        ninv = pow(n, q - 2, q)
        w1n = int(psiinv_tables[l][1]) * ninv % q
        out[l] = [q, 2 * q, mu, ninv, (ninv << 64) // q, w1n, (w1n << 64) // q, (1 << 64) // q, (1 << 64) - q, qbit]
    return out


def p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def ntt(a, n, qs, psi_tables, psiinv_tables, num, division, inverse, barrett, use_tma, group_polys=0, group_stride=0):
    """a: flat uint64 [num*n]; tables [limbs][n]."""
    logn = n.bit_length() - 1
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    tw = np.ascontiguousarray(psiinv_tables if inverse else psi_tables, dtype=np.uint64)
    limbs = len(qs)
    u64p = C.c_ulonglong
    if barrett == 1:
        qv = np.array([int(q) for q in qs], dtype=np.uint64)
        muv = np.array([(1 << (2 * int(q).bit_length())) // int(q) for q in qs], dtype=np.uint64)
        qb = np.array([int(q).bit_length() for q in qs], dtype=np.uint32)
        r = lib().emu_ntt(int(inverse), 1, int(use_tma), logn, p(a, u64p), p(tw, u64p), None, None, p(qv, u64p), p(muv, u64p),
                          p(qb, C.c_uint), num, division, group_polys, C.c_size_t(group_stride))
    else:
        tws = np.ascontiguousarray(np.stack([shoup(tw[l], int(qs[l])) for l in range(limbs)]))
        lc = limb_consts(qs, n, psiinv_tables)
        r = lib().emu_ntt(int(inverse), int(barrett), int(use_tma), logn, p(a, u64p), p(tw, u64p), p(tws, u64p), lc.ctypes.data_as(C.c_void_p),
                          None, None, None, num, division, group_polys, C.c_size_t(group_stride))
    assert r == 0
    return a
