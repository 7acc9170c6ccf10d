"""ctypes front-end of the CPU oracle (oracle/ntt_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (ntt-cuda_b200/) never does.  Every function is a thin numpy wrapper
over the C restatement, which cites the reference file:line it follows.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

u64 = C.c_ulonglong
u64p = C.POINTER(C.c_ulonglong)
u32p = C.POINTER(C.c_uint)
i32p = C.POINTER(C.c_int)
u8p = C.POINTER(C.c_ubyte)


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (a few hundred ms)."""
    src = os.path.join(_HERE, "ntt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_modpow.restype = u64
        _lib.orc_modpow.argtypes = [u64, u64, u64]
        _lib.orc_modinv.restype = u64
        _lib.orc_modinv.argtypes = [u64, u64]
        _lib.orc_bitrev.restype = u64
        _lib.orc_bitrev.argtypes = [u64, C.c_int]
        _lib.orc_qbit.restype = C.c_uint
        _lib.orc_qbit.argtypes = [u64]
        _lib.orc_mu.restype = u64
        _lib.orc_mu.argtypes = [u64, C.c_uint]
        _lib.orc_barrett_mul.restype = u64
        _lib.orc_barrett_mul.argtypes = [u64, u64, u64, u64, C.c_int]
    return _lib


def _p64(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags.c_contiguous
    return a.ctypes.data_as(u64p)


def _p32(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags.c_contiguous
    return a.ctypes.data_as(u32p)


def _pi32(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(i32p)


def _p8(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(u8p)


def A64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


# ------------------------------------------------------------------ number theory / tables
def modpow(a, b, m):
    return int(lib().orc_modpow(a, b, m))


def modinv(a, q):
    return int(lib().orc_modinv(a, q))


def qbit(q):
    return int(lib().orc_qbit(q))


def mu(q, qb=None):
    return int(lib().orc_mu(q, qb if qb is not None else qbit(q)))


def fill_psi_tables(psi, q, n):
    """(psiTable, psiinvTable) as in parameter.h:5-12, psiinv = psi^(q-2) (demo.cu:96-97)."""
    psiinv = modinv(psi, q)
    t = np.empty(n, dtype=np.uint64)
    ti = np.empty(n, dtype=np.uint64)
    lib().orc_fill_psi_tables(u64(psi), u64(q), u64(psiinv), _p64(t), _p64(ti), C.c_uint(n))
    return t, ti


def fill_uniform(count, q, seed):
    a = np.empty(count, dtype=np.uint64)
    lib().orc_fill_uniform(_p64(a), C.c_size_t(count), u64(q), u64(seed))
    return a


# ------------------------------------------------------------------ ring description
@dataclass
class Ring:
    """All limbs of one parameter set + every constant the reference driver derives (demo.cu:62-264)."""
    n: int
    q: list
    psi_roots: list
    t: int = 1024
    gamma: int = 2305843009213683713
    gamma_bits: int = 61
    # derived
    qbit: np.ndarray = field(default=None, repr=False)
    mu: np.ndarray = field(default=None, repr=False)
    psi: np.ndarray = field(default=None, repr=False)      # [r][n]
    psiinv: np.ndarray = field(default=None, repr=False)   # [r][n]

    def __post_init__(self):
        r = len(self.q)
        self.r = r
        self.qa = A64(self.q)
        rp = max(r - 1, 1)
        self.qbit = np.zeros(r, dtype=np.uint32)
        self.mu = np.zeros(r, dtype=np.uint64)
        self.inv_q_last_mod_q = np.zeros(rp, dtype=np.uint64)
        self.qi_div_t = np.zeros(r, dtype=np.uint64)
        self.psiinv_roots = np.zeros(r, dtype=np.uint64)
        self.neg_inv = np.zeros(2, dtype=np.uint64)
        self.prod_t_gamma_mod_q = np.zeros(rp, dtype=np.uint64)
        self.inv_punctured_q = np.zeros(rp, dtype=np.uint64)
        self.bcm = np.zeros(2 * rp, dtype=np.uint64)
        mug = u64(0)
        lib().orc_derive_params(C.c_uint(r), _p64(self.qa), _p64(A64(self.psi_roots)), u64(self.t), u64(self.gamma),
                                C.c_int(self.gamma_bits), _p32(self.qbit), _p64(self.mu), _p64(self.inv_q_last_mod_q),
                                _p64(self.qi_div_t), _p64(self.psiinv_roots), _p64(self.neg_inv),
                                _p64(self.prod_t_gamma_mod_q), C.byref(mug), _p64(self.inv_punctured_q), _p64(self.bcm))
        self.mu_gamma = int(mug.value)
        self.gamma_div_2 = self.gamma >> 1
        self.psi = np.empty((r, self.n), dtype=np.uint64)
        self.psiinv = np.empty((r, self.n), dtype=np.uint64)
        for i in range(r):
            self.psi[i], self.psiinv[i] = fill_psi_tables(int(self.psi_roots[i]), int(self.q[i]), self.n)


# ------------------------------------------------------------------ NTT
def forward_ntt(a, q, psi, mu_=None, qb=None):
    a = A64(a).copy()
    qb = qb if qb is not None else qbit(q)
    mu_ = mu_ if mu_ is not None else mu(q, qb)
    lib().orc_forward_ntt(_p64(a), C.c_uint(a.size), u64(q), u64(mu_), C.c_int(qb), _p64(A64(psi)))
    return a


def inverse_ntt(a, q, psiinv, mu_=None, qb=None):
    a = A64(a).copy()
    qb = qb if qb is not None else qbit(q)
    mu_ = mu_ if mu_ is not None else mu(q, qb)
    lib().orc_inverse_ntt(_p64(a), C.c_uint(a.size), u64(q), u64(mu_), C.c_int(qb), _p64(A64(psiinv)))
    return a


def forward_ntt_fast(a, q, psi):
    a = A64(a).copy()
    lib().orc_forward_ntt_fast(_p64(a), C.c_uint(a.size), u64(q), _p64(A64(psi)))
    return a


def inverse_ntt_fast(a, q, psiinv):
    a = A64(a).copy()
    lib().orc_inverse_ntt_fast(_p64(a), C.c_uint(a.size), u64(q), _p64(A64(psiinv)))
    return a


def forward_ntt_batch(a, n, psi, num, division, q, mu_, qb):
    a = A64(a).copy()
    lib().orc_forward_ntt_batch(_p64(a), C.c_uint(n), _p64(A64(psi)), C.c_uint(num), C.c_uint(division),
                                _p64(A64(q)), _p64(A64(mu_)), _p32(np.ascontiguousarray(qb, dtype=np.uint32)))
    return a


def inverse_ntt_batch(a, n, psiinv, num, division, q, mu_, qb):
    a = A64(a).copy()
    lib().orc_inverse_ntt_batch(_p64(a), C.c_uint(n), _p64(A64(psiinv)), C.c_uint(num), C.c_uint(division),
                                _p64(A64(q)), _p64(A64(mu_)), _p32(np.ascontiguousarray(qb, dtype=np.uint32)))
    return a


def forward_ntt_fast_range(a, n, psi, division, q, p0, p1):
    """In place on `a` (2-D or flat uint64); releases the GIL so it can be threaded."""
    lib().orc_forward_ntt_fast_range(_p64(a), C.c_uint(n), _p64(psi), C.c_uint(division), _p64(q), C.c_uint(p0), C.c_uint(p1))


def inverse_ntt_fast_range(a, n, psiinv, division, q, p0, p1):
    lib().orc_inverse_ntt_fast_range(_p64(a), C.c_uint(n), _p64(psiinv), C.c_uint(division), _p64(q), C.c_uint(p0), C.c_uint(p1))


# ------------------------------------------------------------------ pointwise
def barrett(a, b, q, mu_=None, qb=None):
    a = A64(a).copy()
    qb = qb if qb is not None else qbit(q)
    mu_ = mu_ if mu_ is not None else mu(q, qb)
    lib().orc_barrett(_p64(a), _p64(A64(b)), C.c_uint(a.size), u64(q), u64(mu_), C.c_int(qb))
    return a


def barrett_batch(a, b, n, polys, division, q, mu_, qb):
    a = A64(a).copy()
    lib().orc_barrett_batch(_p64(a), _p64(A64(b)), C.c_uint(n), C.c_uint(polys), C.c_uint(division), _p64(A64(q)),
                            _p64(A64(mu_)), _p32(np.ascontiguousarray(qb, dtype=np.uint32)))
    return a


def barrett_batch_3param(a, b, n, polys, division, q, mu_, qb):
    a = A64(a)
    c = np.empty_like(a)
    lib().orc_barrett_batch_3param(_p64(c), _p64(a), _p64(A64(b)), C.c_uint(n), C.c_uint(polys), C.c_uint(division),
                                   _p64(A64(q)), _p64(A64(mu_)), _p32(np.ascontiguousarray(qb, dtype=np.uint32)))
    return c


def barrett_int(a, b, q, mu_=None, qb=None):
    a = A64(a).copy()
    qb = qb if qb is not None else qbit(q)
    mu_ = mu_ if mu_ is not None else mu(q, qb)
    lib().orc_barrett_int(_p64(a), u64(b), C.c_uint(a.size), u64(q), u64(mu_), C.c_int(qb))
    return a


def mod_t(a, b, t):
    a = A64(a).copy()
    lib().orc_mod_t(_p64(a), u64(b), C.c_uint(a.size), u64(t))
    return a


def poly_add(a, b, q):
    a = A64(a).copy()
    lib().orc_poly_add(_p64(a), _p64(A64(b)), C.c_uint(a.size), u64(q))
    return a


def poly_add_integer(a, b, q):
    a = A64(a).copy()
    lib().orc_poly_add_integer(_p64(a), u64(b), C.c_uint(a.size), u64(q))
    return a


def poly_sub(a, b, q):
    a = A64(a).copy()
    lib().orc_poly_sub(_p64(a), _p64(A64(b)), C.c_uint(a.size), u64(q))
    return a


def poly_negate(a, q):
    a = A64(a).copy()
    lib().orc_poly_negate(_p64(a), C.c_uint(a.size), u64(q))
    return a


def divide_and_round_q_last_inplace_loop(input_poly, last_poly, base_q_i, half_mod, inv_q_last_mod_q_i, mu_, qb):
    a = A64(input_poly).copy()
    lib().orc_divide_and_round_q_last_inplace_loop(_p64(a), _p64(A64(last_poly)), C.c_uint(a.size), u64(base_q_i),
                                                   u64(half_mod), u64(inv_q_last_mod_q_i), u64(mu_), C.c_int(qb))
    return a


def fast_convert(input_poly, n, q_amount, t, gamma, gamma_bits, mu_gamma, bcm):
    """Returns the 2n-element result buffer [mod t | mod gamma] (poly_arithmetic.cuh:265-275)."""
    inp = A64(input_poly)
    res = np.zeros(2 * n, dtype=np.uint64)
    lib().orc_fast_convert_t(_p64(inp), _p64(res), u64(t), _p64(A64(bcm)), C.c_uint(q_amount), C.c_uint(n))
    lib().orc_fast_convert_gamma(_p64(inp), _p64(res), u64(gamma), _p64(A64(bcm)), C.c_uint(q_amount), C.c_int(gamma_bits),
                                 u64(mu_gamma), C.c_uint(n))
    return res


def dec_round(input_poly, n, t, gamma, gamma_div_2):
    inp = A64(input_poly)
    res = np.zeros(n, dtype=np.uint64)
    lib().orc_dec_round(_p64(inp), _p64(res), u64(t), u64(gamma), u64(gamma_div_2), C.c_uint(n))
    return res


def ref_poly_mul(a, b, m):
    a = A64(a)
    d = np.empty_like(a)
    lib().orc_ref_poly_mul(_p64(a), _p64(A64(b)), u64(m), C.c_int(a.size), _p64(d))
    return d


# ------------------------------------------------------------------ sampling
def salsa20_keystream(nbytes, key: bytes, nonce=0):
    out = np.zeros(nbytes, dtype=np.uint8)
    kb = (C.c_ubyte * 32).from_buffer_copy(key)
    lib().orc_salsa20_keystream(_p8(out), C.c_size_t(nbytes), kb, u64(nonce))
    return out


def set_nonce(nonce):
    """Batched pipelines: item k uses Salsa20 nonce k (0 = the reference)."""
    lib().orc_set_nonce(u64(nonce))


def generate_random_default(nbytes):
    out = np.zeros(nbytes, dtype=np.uint8)
    lib().orc_generate_random_default(_p8(out), C.c_uint(nbytes))
    return out


def generate_random(nbytes, prev_key_tail=b"\0" * 8):
    out = np.zeros(nbytes, dtype=np.uint8)
    kb = (C.c_ubyte * 8).from_buffer_copy(prev_key_tail)
    lib().orc_generate_random(_p8(out), C.c_uint(nbytes), kb)
    return out


def convert_ternary(inb, q):
    inb = np.ascontiguousarray(inb, dtype=np.uint8)
    out = np.empty(inb.size, dtype=np.uint64)
    lib().orc_convert_ternary(_p8(inb), _p64(out), C.c_uint(inb.size), u64(q))
    return out


def convert_range(inw, q):
    inw = A64(inw)
    out = np.empty(inw.size, dtype=np.uint64)
    lib().orc_convert_range(_p64(inw), _p64(out), C.c_uint(inw.size), u64(q))
    return out


def gaussian_samples(inw):
    inw = np.ascontiguousarray(inw, dtype=np.uint32)
    out = np.empty(inw.size, dtype=np.int32)
    lib().orc_gaussian_samples(_p32(inw), _pi32(out), C.c_uint(inw.size))
    return out


def convert_gaussian(inw, q):
    inw = np.ascontiguousarray(inw, dtype=np.uint32)
    out = np.empty(inw.size, dtype=np.uint64)
    lib().orc_convert_gaussian(_p32(inw), _p64(out), C.c_uint(inw.size), u64(q))
    return out


def ternary_dist_xq(inb, n, q):
    q = A64(q)
    out = np.empty(n * q.size, dtype=np.uint64)
    lib().orc_ternary_dist_xq(_p8(np.ascontiguousarray(inb, dtype=np.uint8)), _p64(out), C.c_uint(n), C.c_uint(q.size), _p64(q))
    return out


def uniform_dist_xq(inb, n, q):
    q = A64(q)
    out = np.empty(n * q.size, dtype=np.uint64)
    lib().orc_uniform_dist_xq(_p8(np.ascontiguousarray(inb, dtype=np.uint8)), _p64(out), C.c_uint(n), C.c_uint(q.size), _p64(q))
    return out


def gaussian_dist_xq(inb, n, q, samples=None):
    q = A64(q)
    out = np.empty(n * q.size, dtype=np.uint64)
    lib().orc_gaussian_dist_xq(_p8(np.ascontiguousarray(inb, dtype=np.uint8)), _p64(out), C.c_uint(n), C.c_uint(q.size), _p64(q),
                               _pi32(samples))
    return out


def poly_add_negate_xq(a, b, n, q):
    a = A64(a).copy()
    q = A64(q)
    lib().orc_poly_add_negate_xq(_p64(a), _p64(A64(b)), C.c_uint(n), C.c_uint(q.size), _p64(q))
    return a


# ------------------------------------------------------------------ BFV pipelines
def keygen_rns(ring: Ring, e_samples=None):
    """Returns (secret_key[r*n], public_key[2*r*n], temp[r*n], in_bytes)."""
    r, n = ring.r, ring.n
    inb = np.zeros(9 * r * n + 4 * n, dtype=np.uint8)
    sk = np.zeros(r * n, dtype=np.uint64)
    pk = np.zeros(2 * r * n, dtype=np.uint64)
    temp = np.zeros(r * n, dtype=np.uint64)
    lib().orc_keygen_rns(_p8(inb), C.c_uint(r), C.c_uint(n), _p64(ring.qa), _p64(ring.mu), _p32(ring.qbit),
                         _p64(ring.psi), _p64(ring.psiinv), _p64(sk), _p64(pk), _p64(temp), _pi32(e_samples))
    return sk, pk, temp, inb


def encryption_rns(ring: Ring, public_key, m_poly, e0_samples=None, e1_samples=None):
    """Returns (c[2*r*n], e[2*r*n])."""
    r, n = ring.r, ring.n
    c = np.zeros(2 * r * n, dtype=np.uint64)
    e = np.zeros(2 * r * n, dtype=np.uint64)
    inb = np.zeros(9 * n, dtype=np.uint8)
    lib().orc_encryption_rns(_p64(c), _p64(A64(public_key)), _p8(inb), _p64(e), C.c_uint(n), C.c_uint(r), _p64(ring.qa),
                             _p64(ring.mu), _p32(ring.qbit), _p64(ring.inv_q_last_mod_q), _p64(ring.psi), _p64(ring.psiinv),
                             _p64(A64(m_poly)), _p64(ring.qi_div_t), u64(ring.t), _pi32(e0_samples), _pi32(e1_samples))
    return c, e


def decryption_rns(ring: Ring, c, secret_key):
    """Returns (plaintext[n], c_after) -- plaintext is c_after[(rp-1)*n : rp*n] (demo.cu:299)."""
    r, n = ring.r, ring.n
    rp = r - 1
    c = A64(c).copy()
    nb = (C.c_ulonglong * 2)(int(ring.neg_inv[0]), int(ring.neg_inv[1]))
    lib().orc_decryption_rns(_p64(c), _p64(A64(secret_key)), C.c_uint(n), C.c_uint(rp), _p64(ring.qa), _p64(ring.mu),
                             _p32(ring.qbit), _p64(ring.psi), _p64(ring.psiinv), _p64(ring.inv_punctured_q),
                             _p64(ring.prod_t_gamma_mod_q), _p64(ring.bcm), u64(ring.t), u64(ring.gamma), u64(ring.mu_gamma),
                             C.c_int(ring.gamma_bits), nb, u64(ring.gamma_div_2))
    return c[(rp - 1) * n: rp * n].copy(), c
