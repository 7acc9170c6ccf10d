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


def polymul(a, b, n, qs, psi_tables, psiinv_tables, num, division, fwd=True, lazy=True):
    """nttb200_poly_mul_batch (fwd) / nttb200_ntt_domain_mul_inverse_batch (not fwd) on the emulator; returns (result, b after the call)."""
    logn = n.bit_length() - 1
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    b = np.ascontiguousarray(b, dtype=np.uint64).copy()
    psi = np.ascontiguousarray(psi_tables, dtype=np.uint64)
    psiinv = np.ascontiguousarray(psiinv_tables, dtype=np.uint64)
    limbs = len(qs)
    psi_s = np.ascontiguousarray(np.stack([shoup(psi[l], int(qs[l])) for l in range(limbs)]))
    psiinv_s = np.ascontiguousarray(np.stack([shoup(psiinv[l], int(qs[l])) for l in range(limbs)]))
    lc = limb_consts(qs, n, psiinv)
    u = C.c_ulonglong
    r = lib().emu_polymul(int(fwd), int(lazy), logn, p(a, u), p(b, u), p(psi, u), p(psi_s, u), p(psiinv, u), p(psiinv_s, u),
                          lc.ctypes.data_as(C.c_void_p), num, division)
    assert r == 0
    return a, b


# ---- BFV pipelines / pointwise / sampling on the emulator ---------------------------------------------------------------
class EmuRing:
    """Device-side view of an oracle.Ring for the emulator (tables, Shoup companions, LimbConst)."""

    def __init__(self, ring):
        self.ring = ring
        self.psi = np.ascontiguousarray(ring.psi)
        self.psiinv = np.ascontiguousarray(ring.psiinv)
        qs = [int(x) for x in ring.q]
        self.psi_s = np.ascontiguousarray(np.stack([shoup(self.psi[l], qs[l]) for l in range(ring.r)]))
        self.psiinv_s = np.ascontiguousarray(np.stack([shoup(self.psiinv[l], qs[l]) for l in range(ring.r)]))
        self.lc = limb_consts(qs, ring.n, self.psiinv)


def bfv(op, er: EmuRing, barrett, batch=1, nonce0=0, sk=None, pk=None, c=None, m=None, per_item_keys=0):
    """op 0 keygen -> (sk, pk, es); 1 encrypt -> (c, es); 2 decrypt -> (out, c); 3 / 4: encrypt / decrypt through the fused
    NTT (.) key -> INTT kernel with a loaded key (5 / 6: same with the general-modulus policies)."""
    R = er.ring
    n, r = R.n, R.r
    rn = r * n
    u = C.c_ulonglong
    inb = np.zeros(batch * (9 * rn + 4 * n), dtype=np.uint8)
    es = np.zeros(batch * 2 * n, dtype=np.int32)
    sk = np.zeros(batch * rn, dtype=np.uint64) if sk is None else np.ascontiguousarray(sk, dtype=np.uint64).copy()
    pk = np.zeros(batch * 2 * rn, dtype=np.uint64) if pk is None else np.ascontiguousarray(pk, dtype=np.uint64).copy()
    c = np.zeros(batch * 2 * rn, dtype=np.uint64) if c is None else np.ascontiguousarray(c, dtype=np.uint64).copy()
    m = np.zeros(batch * n, dtype=np.uint64) if m is None else np.ascontiguousarray(m, dtype=np.uint64)
    out = np.zeros(batch * n, dtype=np.uint64)
    rc = lib().emu_bfv(op, n, r, p(R.qa, u), p(R.mu, u), p(R.qbit, C.c_uint), p(er.psi, u), p(er.psiinv, u), p(er.psi_s, u), p(er.psiinv_s, u),
                       er.lc.ctypes.data_as(C.c_void_p), int(barrett),
                       p(inb, C.c_ubyte), p(es, C.c_int), p(sk, u), p(pk, u), p(c, u), p(m, u), p(out, u), batch, u(nonce0),
                       p(R.inv_q_last_mod_q, u), p(R.qi_div_t, u), p(R.prod_t_gamma_mod_q, u), p(R.inv_punctured_q, u), p(R.bcm, u),
                       u(R.t), u(R.gamma), u(R.mu_gamma), int(R.gamma_bits), u(int(R.neg_inv[0])), u(int(R.neg_inv[1])), int(per_item_keys))
    assert rc == 0
    if op == 0:
        return sk, pk, es[:batch * n].reshape(batch, n)
    if op in (1, 3, 5):
        return c, es.reshape(batch, 2, n)
    return out.reshape(batch, n), c


def pointwise(op, a, b=None, n=None, s0=0, s1=0, s2=0, i0=0, u0=0, u1=0, qv=None, muv=None, qbitv=None, aux=None, out_size=None):
    u = C.c_ulonglong
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    b = np.zeros(1, dtype=np.uint64) if b is None else np.ascontiguousarray(b, dtype=np.uint64)
    n = a.size if n is None else n
    c = np.zeros(out_size or a.size, dtype=np.uint64)
    z = np.zeros(1, dtype=np.uint64)
    zq = np.zeros(1, dtype=np.uint32)
    qv = z if qv is None else np.ascontiguousarray(qv, dtype=np.uint64)
    muv = z if muv is None else np.ascontiguousarray(muv, dtype=np.uint64)
    qbitv = zq if qbitv is None else np.ascontiguousarray(qbitv, dtype=np.uint32)
    aux = z if aux is None else np.ascontiguousarray(aux, dtype=np.uint64)
    rc = lib().emu_pointwise(op, p(a, u), p(b, u), p(c, u), C.c_size_t(n), u(s0), u(s1), u(s2), int(i0), int(u0), int(u1), p(qv, u), p(muv, u),
                             p(qbitv, C.c_uint), p(aux, u))
    assert rc == 0
    return a, c


def sampling(op, inb, n, u0=0, q=None, nonce=0, stride=0, out_size=None):
    u = C.c_ulonglong
    inb = np.zeros(8, dtype=np.uint8) if inb is None else np.ascontiguousarray(inb).view(np.uint8)
    out = np.zeros(out_size if out_size else n, dtype=np.uint64)
    q = np.zeros(1, dtype=np.uint64) if q is None else np.ascontiguousarray(q, dtype=np.uint64)
    rc = lib().emu_sampling(op, p(inb, C.c_ubyte), p(out, u), None, C.c_size_t(n), int(u0), p(q, u), u(nonce), C.c_size_t(stride))
    assert rc == 0
    return out


class EmuBfv:
    """Adapter with the nttb200.Bfv decrypt_partial / decrypt_finish interface, backed by the emulator (CPU tests of the
    multi-GPU control flow)."""

    def __init__(self, er: EmuRing):
        self.er, self.n, self.r = er, er.ring.n, er.ring.r

    def _call(self, op, c_shard, sk_shard, part, out, batch, first, count, half=0):
        R, er, u = self.er.ring, self.er, C.c_ulonglong
        z = np.zeros(1, dtype=np.uint64)
        rc = lib().emu_bfv_sharded(op, R.n, R.r, p(R.qa, u), p(R.mu, u), p(R.qbit, C.c_uint), p(er.psi, u), p(er.psiinv, u), p(er.psi_s, u),
                                   p(er.psiinv_s, u), er.lc.ctypes.data_as(C.c_void_p), p(c_shard if c_shard is not None else z, u),
                                   p(sk_shard if sk_shard is not None else z, u), p(part, u), p(out if out is not None else z, u), batch, first,
                                   count, p(R.prod_t_gamma_mod_q, u), p(R.inv_punctured_q, u), p(R.bcm, u), u(R.t), u(R.gamma), u(R.mu_gamma),
                                   int(R.gamma_bits), u(int(R.neg_inv[0])), u(int(R.neg_inv[1])), int(half))
        assert rc == 0

    def decrypt_partial(self, partial, c_shard, sk_shard, first, count, batch=1, sk_per_item=False, shard_half_limbs=0):
        assert not sk_per_item
        self._call(0, c_shard, sk_shard, partial, None, batch, first, count, shard_half_limbs)

    def encrypt(self, c, pk, m, batch=1, nonce0=0):
        """nttb200.Bfv.encrypt interface on the emulator (in place into the numpy array c)."""
        out, _ = bfv(1, self.er, 0, batch=batch, nonce0=nonce0, pk=pk, m=m)
        c[:] = out

    def decrypt_finish(self, m_out, partial_sum, batch=1):
        self._call(1, None, None, partial_sum, m_out, batch, 0, 0)


# ---- building blocks of the fused-epilogue encryption / fused sharded decryption (mirror csrc/bfv.cu enc_* / dec_*) --------------------
def enc_epi_limbs(R):
    """EncEpiLimb[r-1] exactly as nttb200_bfv_create builds it (csrc/bfv.cu)."""
    r = R.r
    qs = [int(x) for x in R.q]
    out = np.zeros((r - 1, 8), dtype=np.uint64)
    for i in range(r - 1):
        q, iql = qs[i], int(R.inv_q_last_mod_q[i])
        out[i] = [q, 2 * q, iql, (iql << 64) // q, int(R.qi_div_t[i]), (qs[r - 1] >> 1) % q + 3 * q, ((1 << 64) - 1) // q, int(qs[r - 1] > 2 * q)]
    assert lib().emu_sizeof_encepilimb() == 64
    return out


class EmuBlocks:
    """enc_sample / enc_front / enc_finish_last / enc_finish_limbs / dec_partial / dec_finish / dec_expand16 on the emulator, with a key
    pair loaded (companions built here, as nttb200_bfv_load_keys does)."""

    def __init__(self, er: EmuRing, sk=None, pk=None):
        self.er, self.R = er, er.ring
        R = self.R
        self.n, self.r = R.n, R.r
        qs = [int(x) for x in R.q]
        self.K = enc_epi_limbs(R)
        self.tsh = int(R.t).bit_length() - 1
        self.sk = self.sk_s = self.pk = self.pk_s = None
        self.fused = False
        if sk is not None:
            self.sk = np.ascontiguousarray(sk, dtype=np.uint64)
            self.sk_s = np.concatenate([shoup(self.sk[l * R.n:(l + 1) * R.n], qs[l]) for l in range(R.r)])
        if pk is not None:
            self.pk = np.ascontiguousarray(pk, dtype=np.uint64)
            self.pk_s = np.concatenate([shoup(self.pk[i * R.n:(i + 1) * R.n], qs[i % R.r]) for i in range(2 * R.r)])

    def _ring(self):
        R, er, u = self.R, self.er, C.c_ulonglong
        return (R.n, R.r, p(R.qa, u), p(R.mu, u), p(R.qbit, C.c_uint), p(er.psi, u), p(er.psiinv, u), p(er.psi_s, u), p(er.psiinv_s, u),
                er.lc.ctypes.data_as(C.c_void_p))

    def _enc(self, op, first=0, count=0, slots=0, items=0, c=None, ub=None, es8=None, cl=None, cl_is=0, cl_hs=0, m=None, nonce0=0, i0=0, i1=0):
        u = C.c_ulonglong
        z = np.zeros(1, dtype=np.uint64)
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None   # noqa: E731
        rc = lib().emu_enc_blocks(op, *self._ring(), first, count, slots, items, vp(c), vp(ub), vp(es8),
                                  p(self.pk if self.pk is not None else z, u), p(self.pk_s if self.pk_s is not None else z, u), vp(cl),
                                  C.c_size_t(cl_is), C.c_size_t(cl_hs), vp(m), self.K.ctypes.data_as(C.c_void_p), u(self.R.t), self.tsh,
                                  u(nonce0), i0, i1)
        assert rc == 0, rc

    def enc_sample(self, ub, es8, items, nonce0, want_u, want_e):
        self._enc(0, items=items, ub=ub, es8=es8, nonce0=nonce0, i0=int(want_u), i1=int(want_e))

    def enc_front(self, c, slots, first, count, items, ub):
        self._enc(1, first, count, slots, items, c=c, ub=ub)

    def enc_finish_last(self, cl, cl_is, cl_hs, es8, items):
        self._enc(2, items=items, cl=cl, cl_is=cl_is, cl_hs=cl_hs, es8=es8)

    def enc_finish_limbs(self, c, slots, first, count, items, cl, cl_is, cl_hs, es8, m, fused=None):
        """fused: epilogue in the store of the last inverse kernel (A/B variant); default (the product's): plain pass + epilogue kernel"""
        R, u = self.R, C.c_ulonglong
        fused = self.fused if fused is None else fused
        qs = [int(x) for x in R.q]
        lazy = all(qs[-1] <= 2 * q for q in qs[:-1])
        lib().emu_set_enc_arrays(p(R.inv_q_last_mod_q, u), p(R.qi_div_t, u))
        self._enc(3, first, count, slots, items, c=c, cl=cl, cl_is=cl_is, cl_hs=cl_hs, es8=es8, m=m, i0=int(fused), i1=int(lazy))

    def _dec(self, op, first=0, count=0, slots=0, items=0, c=None, part=None, out=None, packed=0, out16=0):
        R, u = self.R, C.c_ulonglong
        z = np.zeros(1, dtype=np.uint64)
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None   # noqa: E731
        rc = lib().emu_dec_blocks(op, *self._ring(), first, count, slots, items, vp(c), p(self.sk if self.sk is not None else z, u),
                                  p(self.sk_s if self.sk_s is not None else z, u), vp(part), vp(out), int(packed), int(out16),
                                  p(R.prod_t_gamma_mod_q, u), p(R.inv_punctured_q, u), p(R.bcm, u), u(R.t), u(R.gamma), u(R.mu_gamma),
                                  int(R.gamma_bits), u(int(R.neg_inv[0])), u(int(R.neg_inv[1])))
        assert rc == 0, rc

    def dec_partial(self, part, packed, c_shard, slots, first, count, items):
        self._dec(0, first, count, slots, items, c=c_shard, part=part, packed=packed)

    def dec_finish(self, out, out16, part, packed, items):
        self._dec(1, items=items, part=part, out=out, packed=packed, out16=out16)

    def dec_expand16(self, plain16, out, items):
        self._dec(2, items=items, part=plain16, out=out)
