"""nttb200 -- Python host side over the C ABI of libnttb200.so (include/nttb200.h).

PyTorch is used only for device memory and streams; every operation is a hand-written sm_100a kernel inside
libnttb200.so.  There is NO CPU fallback: importing works without the library (so CPU-only tooling can read
`params`), but any compute call raises if the shared library or a CUDA device is missing.

Function names mirror the reference's host API (ntt_60bit.cuh, poly_arithmetic.cuh, distributions.cuh,
bfv_*.cuh) so the parity tests read like the reference's own drivers.
"""
from __future__ import annotations

import ctypes as C
import os

from . import params  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NTTB200_LIB") or os.path.join(_HERE, "libnttb200.so")   # NTTB200_LIB: tuning variants

_lib = None
u64 = C.c_ulonglong
vp = C.c_void_p


class NttB200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads libnttb200.so; raises loudly if it has not been built (python ntt-cuda_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NttB200Error(f"{LIB_PATH} is missing: build it with `python ntt-cuda_b200/build.py` (no CPU fallback exists)")
        _lib = C.CDLL(LIB_PATH)
        _lib.nttb200_error_string.restype = C.c_char_p
        _lib.nttb200_error_string.argtypes = [C.c_int]
    return _lib


def check(code: int):
    if code != 0:
        raise NttB200Error(f"nttb200 error {code}: {lib().nttb200_error_string(code).decode()}")


def ptr(x) -> int:
    """Device/host address of a torch tensor, numpy array or int."""
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(type(x))


def download(dev_ptr: int, count: int, dtype="uint64"):
    """Synchronous device -> host copy of `count` elements into a new numpy array."""
    import numpy as np
    out = np.empty(count, dtype=dtype)
    check(lib().nttb200_download(vp(out.ctypes.data), vp(ptr(dev_ptr)), C.c_size_t(out.nbytes)))
    return out


def _stream(stream) -> int:
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


def _arr64(vals):
    vals = [int(v) for v in vals]
    return (u64 * len(vals))(*vals)


class Context:
    """Per-ring state in HBM: reference-layout psi / psiinv tables [limbs][n], Shoup companions, per-limb constants.
    Replaces the reference drivers' manual set-up (demo.cu:62-196)."""

    def __init__(self, n: int, q, psi_roots=None, psi_tables=None, psiinv_tables=None):
        self.n, self.q, self.limbs = int(n), [int(v) for v in q], len(q)
        self._h = vp()
        if psi_tables is not None:
            import numpy as np
            pt = np.ascontiguousarray(psi_tables, dtype=np.uint64)
            pit = np.ascontiguousarray(psiinv_tables, dtype=np.uint64)
            check(lib().nttb200_ctx_create_from_tables(C.byref(self._h), C.c_uint(self.n), C.c_uint(self.limbs), _arr64(self.q),
                                                       vp(pt.ctypes.data), vp(pit.ctypes.data)))
        else:
            check(lib().nttb200_ctx_create(C.byref(self._h), C.c_uint(self.n), C.c_uint(self.limbs), _arr64(self.q), _arr64(psi_roots)))
        a, b = vp(), vp()
        check(lib().nttb200_ctx_tables(self._h, C.byref(a), C.byref(b)))
        self.psi_table, self.psiinv_table = a.value, b.value            # device addresses, reference layout
        qd, md, bd = vp(), vp(), vp()
        check(lib().nttb200_ctx_consts(self._h, C.byref(qd), C.byref(md), C.byref(bd)))
        self.q_dev, self.mu_dev, self.qbit_dev = qd.value, md.value, bd.value

    def close(self):
        if self._h:
            lib().nttb200_ctx_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tma(self, enable: bool):
        check(lib().nttb200_ctx_set_tma(self._h, C.c_int(1 if enable else 0)))

    # forwardNTT_batch / inverseNTT_batch (ntt_60bit.cuh:608, :652), fast path
    def forward_ntt_batch(self, a, num: int, division: int, stream=None):
        check(lib().nttb200_forward_ntt_batch(self._h, vp(ptr(a)), C.c_uint(num), C.c_uint(division), vp(_stream(stream))))

    def inverse_ntt_batch(self, a, num: int, division: int, stream=None):
        check(lib().nttb200_inverse_ntt_batch(self._h, vp(ptr(a)), C.c_uint(num), C.c_uint(division), vp(_stream(stream))))

    def poly_mul_batch(self, a, b, num: int, division: int, stream=None):
        """a <- a * b mod (X^n + 1, q_limb), fused (full_poly_mul_device, poly_arithmetic.cuh:296); b is clobbered."""
        check(lib().nttb200_poly_mul_batch(self._h, vp(ptr(a)), vp(ptr(b)), C.c_uint(num), C.c_uint(division), vp(_stream(stream))))

    def ntt_domain_mul_inverse_batch(self, a, b, num: int, division: int, stream=None):
        """a <- INTT(a (.) b), both operands in the NTT domain; b is only read."""
        check(lib().nttb200_ntt_domain_mul_inverse_batch(self._h, vp(ptr(a)), vp(ptr(b)), C.c_uint(num), C.c_uint(division),
                                                         vp(_stream(stream))))

    def ntt_pass(self, a, num: int, division: int, inverse: bool, which: int, stream=None):
        """Profiling hook: only the first / second kernel of the transform (execution order)."""
        check(lib().nttb200_ntt_pass(self._h, vp(ptr(a)), C.c_uint(num), C.c_uint(division), C.c_int(int(inverse)), C.c_int(which),
                                     vp(_stream(stream))))

    # host-buffer (end-to-end) variants: numpy uint64 arrays (pinned or pageable); synchronous
    def forward_ntt_batch_host(self, a_in, a_out, num: int, division: int):
        check(lib().nttb200_forward_ntt_batch_host(self._h, vp(ptr(a_in)), vp(ptr(a_out)), C.c_uint(num), C.c_uint(division)))

    def inverse_ntt_batch_host(self, a_in, a_out, num: int, division: int):
        check(lib().nttb200_inverse_ntt_batch_host(self._h, vp(ptr(a_in)), vp(ptr(a_out)), C.c_uint(num), C.c_uint(division)))

    # compact wire format (n * qbit bits per polynomial; include/nttb200.h): sizes, device <-> device conversion, host-buffer transforms
    def packed_words(self, num: int, division: int) -> int:
        w = C.c_size_t()
        check(lib().nttb200_polys_packed_words(self._h, C.c_uint(num), C.c_uint(division), C.byref(w)))
        return int(w.value)

    def pack_polys(self, packed, a, num: int, division: int, stream=None):
        check(lib().nttb200_pack_polys(self._h, vp(ptr(packed)), vp(ptr(a)), C.c_uint(num), C.c_uint(division), vp(_stream(stream))))

    def unpack_polys(self, a, packed, num: int, division: int, stream=None):
        check(lib().nttb200_unpack_polys(self._h, vp(ptr(a)), vp(ptr(packed)), C.c_uint(num), C.c_uint(division), vp(_stream(stream))))

    def forward_ntt_batch_host_packed(self, p_in, p_out, num: int, division: int):
        check(lib().nttb200_forward_ntt_batch_host_packed(self._h, vp(ptr(p_in)), vp(ptr(p_out)), C.c_uint(num), C.c_uint(division)))

    def inverse_ntt_batch_host_packed(self, p_in, p_out, num: int, division: int):
        check(lib().nttb200_inverse_ntt_batch_host_packed(self._h, vp(ptr(p_in)), vp(ptr(p_out)), C.c_uint(num), C.c_uint(division)))


# ---- stateless reference-contract entry points (what include/dropin/ntt_60bit.cuh forwards to) ----------------------
def forwardNTT_batch(device_a, n, psi_powers, num, division, q_cons, mu_cons, q_bit_cons, stream=None):
    """ntt_60bit.cuh:608; q_cons / mu_cons / q_bit_cons are device arrays (the reference's __constant__ symbols)."""
    check(lib().nttb200_ref_forward_ntt_batch(vp(ptr(device_a)), C.c_uint(n), vp(ptr(psi_powers)), C.c_uint(num), C.c_uint(division),
                                              vp(ptr(q_cons)), vp(ptr(mu_cons)), vp(ptr(q_bit_cons)), vp(_stream(stream))))


def inverseNTT_batch(device_a, n, psiinv_powers, num, division, q_cons, mu_cons, q_bit_cons, stream=None):
    """ntt_60bit.cuh:652"""
    check(lib().nttb200_ref_inverse_ntt_batch(vp(ptr(device_a)), C.c_uint(n), vp(ptr(psiinv_powers)), C.c_uint(num), C.c_uint(division),
                                              vp(ptr(q_cons)), vp(ptr(mu_cons)), vp(ptr(q_bit_cons)), vp(_stream(stream))))


def forwardNTT(device_a, n, stream, q, mu, bit_length, psi_powers):
    """ntt_60bit.cuh:314"""
    check(lib().nttb200_ref_forward_ntt(vp(ptr(device_a)), C.c_uint(n), vp(_stream(stream)), u64(q), u64(mu), C.c_int(bit_length),
                                        vp(ptr(psi_powers))))


def inverseNTT(device_a, n, stream, q, mu, bit_length, psiinv_powers):
    """ntt_60bit.cuh:350"""
    check(lib().nttb200_ref_inverse_ntt(vp(ptr(device_a)), C.c_uint(n), vp(_stream(stream)), u64(q), u64(mu), C.c_int(bit_length),
                                        vp(ptr(psiinv_powers))))


# ---- coefficient-wise kernels (poly_arithmetic.cuh) ------------------------------------------------------------------------
def _call(name, *args):
    check(getattr(lib(), name)(*args))


def barrett(a, b, n, q, mu, qbit, stream=None):
    """barrett<<<n/256,256>>>(a, b, q, mu, qbit), poly_arithmetic.cuh:9"""
    _call("nttb200_barrett", vp(ptr(a)), vp(ptr(b)), C.c_uint(n), u64(q), u64(mu), C.c_int(qbit), vp(_stream(stream)))


def barrett_batch(a, b, n, polys, division, q_cons, mu_cons, q_bit_cons, stream=None):
    _call("nttb200_barrett_batch", vp(ptr(a)), vp(ptr(b)), C.c_uint(n), C.c_uint(polys), C.c_uint(division), vp(ptr(q_cons)), vp(ptr(mu_cons)),
          vp(ptr(q_bit_cons)), vp(_stream(stream)))


def barrett_batch_3param(c, a, b, n, polys, division, q_cons, mu_cons, q_bit_cons, stream=None):
    _call("nttb200_barrett_batch_3param", vp(ptr(c)), vp(ptr(a)), vp(ptr(b)), C.c_uint(n), C.c_uint(polys), C.c_uint(division), vp(ptr(q_cons)),
          vp(ptr(mu_cons)), vp(ptr(q_bit_cons)), vp(_stream(stream)))


def poly_mul_int(a, b, n, stream, q, mu, bit_length):
    _call("nttb200_barrett_int", vp(ptr(a)), u64(b), C.c_uint(n), u64(q), u64(mu), C.c_int(bit_length), vp(_stream(stream)))


def poly_mul_int_t(a, b, n, stream, t):
    _call("nttb200_mod_t", vp(ptr(a)), u64(b), C.c_uint(n), u64(t), vp(_stream(stream)))


def poly_add_device(a, b, n, stream, q):
    _call("nttb200_poly_add", vp(ptr(a)), vp(ptr(b)), C.c_uint(n), u64(q), vp(_stream(stream)))


def poly_add_integer_device(a, b, n, stream, q):
    _call("nttb200_poly_add_integer", vp(ptr(a)), u64(b), C.c_uint(n), u64(q), vp(_stream(stream)))


def poly_sub_device(a, b, n, stream, q):
    _call("nttb200_poly_sub", vp(ptr(a)), vp(ptr(b)), C.c_uint(n), u64(q), vp(_stream(stream)))


def poly_negate_device(a, n, stream, q):
    _call("nttb200_poly_negate", vp(ptr(a)), C.c_uint(n), u64(q), vp(_stream(stream)))


def divide_and_round_q_last_inplace_loop(input_poly, rns_poly_minus1, n, base_q_i, half_mod, inv_q_last_mod_q_i, mu, qbit, stream=None):
    _call("nttb200_divide_and_round_q_last_inplace_loop", vp(ptr(input_poly)), vp(ptr(rns_poly_minus1)), C.c_uint(n), u64(base_q_i), u64(half_mod),
          u64(inv_q_last_mod_q_i), u64(mu), C.c_int(qbit), vp(_stream(stream)))


def fast_convert_array_kernels(input_poly, result_poly, t, bcm, q_amount, gamma, gamma_bits, mu_gamma, n, stream=None):
    _call("nttb200_fast_convert_array", vp(ptr(input_poly)), vp(ptr(result_poly)), u64(t), vp(ptr(bcm)), C.c_uint(q_amount), u64(gamma),
          C.c_int(gamma_bits), u64(mu_gamma), C.c_uint(n), vp(_stream(stream)))


def dec_round(input_poly, result_poly, t, gamma, gamma_div_2, n, stream=None):
    _call("nttb200_dec_round", vp(ptr(input_poly)), vp(ptr(result_poly)), u64(t), u64(gamma), u64(gamma_div_2), C.c_uint(n), vp(_stream(stream)))


# ---- sampling (distributions.cuh) ---------------------------------------------------------------------------------------
def generate_random_default(a, nbytes, stream=None):
    _call("nttb200_generate_random_default", vp(ptr(a)), C.c_uint(nbytes), vp(_stream(stream)))


def generate_random(a, nbytes, stream=None):
    _call("nttb200_generate_random", vp(ptr(a)), C.c_uint(nbytes), vp(_stream(stream)))


def salsa20_keystream(out, blocks_per_stream, streams, stream_stride, key: bytes, nonce0, stream=None):
    kb = (C.c_ubyte * 32).from_buffer_copy(key)
    _call("nttb200_salsa20_keystream", vp(ptr(out)), u64(blocks_per_stream), u64(streams), C.c_size_t(stream_stride), kb, u64(nonce0),
          vp(_stream(stream)))


def gaussian_dist(inp, out, n, stream, q):
    _call("nttb200_gaussian_dist", vp(ptr(inp)), vp(ptr(out)), C.c_uint(n), vp(_stream(stream)), u64(q))


def uniform_dist(inp, out, n, stream, q):
    _call("nttb200_uniform_dist", vp(ptr(inp)), vp(ptr(out)), C.c_uint(n), vp(_stream(stream)), u64(q))


def ternary_dist(inp, out, n, stream, q):
    _call("nttb200_ternary_dist", vp(ptr(inp)), vp(ptr(out)), C.c_uint(n), vp(_stream(stream)), u64(q))


def ternary_dist_xq(inp, sk, n, q_amount, q_cons, stream=None):
    _call("nttb200_ternary_dist_xq", vp(ptr(inp)), vp(ptr(sk)), C.c_uint(n), C.c_uint(q_amount), vp(ptr(q_cons)), vp(_stream(stream)))


def uniform_dist_xq(inp, pk, n, q_amount, q_cons, stream=None):
    _call("nttb200_uniform_dist_xq", vp(ptr(inp)), vp(ptr(pk)), C.c_uint(n), C.c_uint(q_amount), vp(ptr(q_cons)), vp(_stream(stream)))


def gaussian_dist_xq(inp, temp, n, q_amount, q_cons, stream=None):
    _call("nttb200_gaussian_dist_xq", vp(ptr(inp)), vp(ptr(temp)), C.c_uint(n), C.c_uint(q_amount), vp(ptr(q_cons)), vp(_stream(stream)))


def poly_add_negate_xq(a, b, n, q_amount, q_cons, stream=None):
    _call("nttb200_poly_add_negate_xq", vp(ptr(a)), vp(ptr(b)), C.c_uint(n), C.c_uint(q_amount), vp(ptr(q_cons)), vp(_stream(stream)))


def convert_ternary_gaussian_x2(inp, c, e, n, q_amount, q_cons, stream=None):
    _call("nttb200_convert_ternary_gaussian_x2", vp(ptr(inp)), vp(ptr(c)), vp(ptr(e)), C.c_uint(n), C.c_uint(q_amount), vp(ptr(q_cons)),
          vp(_stream(stream)))


# ---- BFV pipelines -------------------------------------------------------------------------------------------------------
class Bfv:
    """Batched BFV keygen / encrypt / decrypt on one ring (replaces the per-driver set-up of demo.cu:62-272 and the
    single-item keygen_rns / encryption_rns / decryption_rns calls).  Layouts are the reference's."""

    def __init__(self, n, q, psi_roots, t=params.T, gamma=params.GAMMA):
        self.n, self.q, self.r, self.t, self.gamma = int(n), [int(v) for v in q], len(q), int(t), int(gamma)
        self._h = vp()
        check(lib().nttb200_bfv_create(C.byref(self._h), C.c_uint(self.n), C.c_uint(self.r), _arr64(self.q), _arr64(psi_roots), u64(t), u64(gamma)))

    def close(self):
        if self._h:
            lib().nttb200_bfv_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reserve(self, batch):
        check(lib().nttb200_bfv_reserve(self._h, C.c_uint(batch)))

    def load_keys(self, sk=None, pk=None, stream=None):
        """Keep a key pair (and its Shoup companions) in the context; encrypt(pk=None) / decrypt(sk=None) then run the
        fused NTT (.) key -> INTT path."""
        check(lib().nttb200_bfv_load_keys(self._h, vp(ptr(sk)), vp(ptr(pk)), vp(_stream(stream))))

    def keygen(self, sk, pk, batch=1, nonce0=0, stream=None):
        """sk[batch][r][n], pk[batch][2][r][n] (bfv_keygen.cuh:95)"""
        check(lib().nttb200_bfv_keygen(self._h, vp(ptr(sk)), vp(ptr(pk)), C.c_uint(batch), u64(nonce0), vp(_stream(stream))))

    def encrypt(self, c, pk, m, batch=1, nonce0=0, pk_per_item=False, stream=None):
        """c[batch][2][r][n], m[batch][n] (bfv_encryption.cuh:223)"""
        check(lib().nttb200_bfv_encrypt(self._h, vp(ptr(c)), vp(ptr(pk)), C.c_int(int(pk_per_item)), vp(ptr(m)), C.c_uint(batch), u64(nonce0),
                                        vp(_stream(stream))))

    def decrypt(self, m_out, c, sk, batch=1, sk_per_item=False, stream=None):
        """m_out[batch][n] (bfv_decryption.cuh:76)"""
        check(lib().nttb200_bfv_decrypt(self._h, vp(ptr(m_out)), vp(ptr(c)), vp(ptr(sk)), C.c_int(int(sk_per_item)), C.c_uint(batch),
                                        vp(_stream(stream))))


    def packed_words(self) -> int:
        """64-bit words of one ciphertext in the compact wire format."""
        lib().nttb200_bfv_packed_words.restype = C.c_size_t
        return int(lib().nttb200_bfv_packed_words(self._h))

    def pack(self, packed, c, batch=1, stream=None):
        check(lib().nttb200_bfv_pack(self._h, vp(ptr(packed)), vp(ptr(c)), C.c_uint(batch), vp(_stream(stream))))

    def unpack(self, c, packed, batch=1, stream=None):
        check(lib().nttb200_bfv_unpack(self._h, vp(ptr(c)), vp(ptr(packed)), C.c_uint(batch), vp(_stream(stream))))

    def key_packed_words(self, polys) -> int:
        lib().nttb200_bfv_key_packed_words.restype = C.c_size_t
        return int(lib().nttb200_bfv_key_packed_words(self._h, C.c_uint(polys)))

    def pack_key(self, packed, key, polys, batch=1, stream=None):
        """sk[batch][r][n] (polys = r) / pk[batch][2][r][n] (polys = 2r) -> bit-packed words, all r limbs."""
        check(lib().nttb200_bfv_pack_key(self._h, vp(ptr(packed)), vp(ptr(key)), C.c_uint(polys), C.c_uint(batch), vp(_stream(stream))))

    def unpack_key(self, key, packed, polys, batch=1, stream=None):
        check(lib().nttb200_bfv_unpack_key(self._h, vp(ptr(key)), vp(ptr(packed)), C.c_uint(polys), C.c_uint(batch), vp(_stream(stream))))

    def pack_host(self, packed_host, c, batch=1, stream=None):
        """device ciphertexts -> packed numpy / pinned host buffer (synchronous)"""
        check(lib().nttb200_bfv_pack_host(self._h, vp(ptr(packed_host)), vp(ptr(c)), C.c_uint(batch), vp(_stream(stream))))

    def unpack_host(self, c, packed_host, batch=1, stream=None):
        check(lib().nttb200_bfv_unpack_host(self._h, vp(ptr(c)), vp(ptr(packed_host)), C.c_uint(batch), vp(_stream(stream))))

    def encrypt_host(self, c_host, m_host, batch, nonce0=0, packed=True):
        """m_host[batch][n] -> ciphertexts in host memory (wire format when packed); loaded public key; synchronous."""
        check(lib().nttb200_bfv_encrypt_host(self._h, vp(ptr(c_host)), C.c_int(int(packed)), vp(ptr(m_host)), C.c_uint(batch), u64(nonce0)))

    def decrypt_host(self, m_host, c_host, batch, packed=True):
        check(lib().nttb200_bfv_decrypt_host(self._h, vp(ptr(m_host)), vp(ptr(c_host)), C.c_int(int(packed)), C.c_uint(batch)))

    # ---- ciphertext x ciphertext multiplication + relinearisation -------------------------------------------------
    def relin_keygen(self, sk, nonce0=1 << 62, stream=None):
        """evk from sk[r][n] (NTT domain); digit i samples nonce nonce0 + i (default: a nonce range keygen / encrypt do not use)."""
        check(lib().nttb200_bfv_relin_keygen(self._h, vp(ptr(sk)), u64(nonce0), vp(_stream(stream))))

    def relin_key(self):
        """(device address, words) of evk[r-1][2][r-1][n]"""
        a, w = vp(), C.c_size_t(0)
        check(lib().nttb200_bfv_relin_key(self._h, C.byref(a), C.byref(w)))
        return a.value, int(w.value)

    def mul_tensor(self, y, c_a, c_b, batch=1, stream=None):
        """y[batch][3][r-1][n] = round(t/Q * (c_a (x) c_b)): degree-2 ciphertext (coefficient domain)."""
        check(lib().nttb200_bfv_mul_tensor(self._h, vp(ptr(y)), vp(ptr(c_a)), vp(ptr(c_b)), C.c_uint(batch), vp(_stream(stream))))

    def relinearize(self, c_out, y, batch=1, stream=None):
        check(lib().nttb200_bfv_relinearize(self._h, vp(ptr(c_out)), vp(ptr(y)), C.c_uint(batch), vp(_stream(stream))))

    def mul(self, c_out, c_a, c_b, batch=1, stream=None):
        """c_out <- relin(c_a * c_b): Dec = m_a * m_b mod (X^n + 1, t)."""
        check(lib().nttb200_bfv_mul(self._h, vp(ptr(c_out)), vp(ptr(c_a)), vp(ptr(c_b)), C.c_uint(batch), vp(_stream(stream))))

    def mul_aux_base(self):
        cnt = C.c_uint(0)
        buf = (u64 * 64)()
        check(lib().nttb200_bfv_mul_aux_base(self._h, buf, C.byref(cnt)))
        return [int(buf[i]) for i in range(cnt.value)]

    def add(self, c_a, c_b, batch=1, stream=None):
        """c_a <- c_a + c_b (homomorphic addition: Dec = m_a + m_b mod t)."""
        check(lib().nttb200_bfv_add(self._h, vp(ptr(c_a)), vp(ptr(c_b)), C.c_uint(batch), vp(_stream(stream))))

    def add_plain(self, c, m_poly, batch=1, plain_per_item=False, stream=None):
        """c <- c + m for a plaintext polynomial m[n] (or m[batch][n]): Dec = m_c + m mod t."""
        check(lib().nttb200_bfv_add_plain(self._h, vp(ptr(c)), vp(ptr(m_poly)), C.c_int(int(plain_per_item)), C.c_uint(batch),
                                          vp(_stream(stream))))

    def mul_plain(self, c, p_poly, batch=1, plain_per_item=False, stream=None):
        """c <- c * p for a plaintext polynomial p[n] (or p[batch][n]): Dec = m * p mod (X^n + 1, t)."""
        check(lib().nttb200_bfv_mul_plain(self._h, vp(ptr(c)), vp(ptr(p_poly)), C.c_int(int(plain_per_item)), C.c_uint(batch),
                                          vp(_stream(stream))))

    def decrypt_partial(self, partial, c_shard, sk_shard, first_limb, limb_count, batch=1, sk_per_item=False, shard_half_limbs=0,
                        stream=None):
        """Limb-sharded decryption, this GPU's share: partial[batch][2][n] (all-reduce SUM it, then decrypt_finish).
        c_shard[batch][2][shard_half_limbs][n]; shard_half_limbs = 0 means compact (= limb_count)."""
        check(lib().nttb200_bfv_decrypt_partial(self._h, vp(ptr(partial)), vp(ptr(c_shard)), vp(ptr(sk_shard)), C.c_int(int(sk_per_item)),
                                                C.c_uint(first_limb), C.c_uint(limb_count), C.c_uint(shard_half_limbs), C.c_uint(batch),
                                                vp(_stream(stream))))

    def decrypt_finish(self, m_out, partial_sum, batch=1, stream=None):
        check(lib().nttb200_bfv_decrypt_finish(self._h, vp(ptr(m_out)), vp(ptr(partial_sum)), C.c_uint(batch), vp(_stream(stream))))

    # ---- round 2: sampling key, fused-epilogue knob, limb-sharded calls with the collectives inside the library -------------------
    def set_sampling_key(self, key: bytes):
        """32-byte Salsa20 key of keygen / encrypt (default: the reference's 32 x 0x01 -- parity tests only)."""
        assert len(key) == 32
        check(lib().nttb200_bfv_set_sampling_key(self._h, (C.c_ubyte * 32).from_buffer_copy(key)))

    def set_fused_epilogue(self, enable: bool):
        check(lib().nttb200_bfv_set_fused_epilogue(self._h, C.c_int(int(enable))))

    def shard_config(self, mode=4, chunks=0):
        check(lib().nttb200_bfv_shard_config(self._h, C.c_int(mode), C.c_uint(chunks)))

    def shard_words(self, comm, batch):
        return shard_plan(self.r - 1, self.n, batch, comm.world, comm.rank)[1]

    def encrypt_sharded(self, comm, c_shard, m, batch, nonce0=0, stream=None):
        """Limb-sharded encryption_rns over comm's ranks (loaded public key); c_shard: this rank's tiles (shard_plan)."""
        check(lib().nttb200_bfv_encrypt_sharded(self._h, comm._h, vp(ptr(c_shard)), vp(ptr(m)), C.c_uint(batch), u64(nonce0), vp(_stream(stream))))

    def decrypt_sharded(self, comm, m_out, c_shard, batch, stream=None):
        """Limb-sharded decryption_rns (loaded secret key); m_out[batch][n] complete on every rank; c_shard is consumed."""
        check(lib().nttb200_bfv_decrypt_sharded(self._h, comm._h, vp(ptr(m_out)), vp(ptr(c_shard)), C.c_uint(batch), vp(_stream(stream))))

    def shard_from_full(self, world, rank, c_shard, c_full, batch, stream=None):
        check(lib().nttb200_bfv_shard_from_full(self._h, C.c_uint(world), C.c_uint(rank), vp(ptr(c_shard)), vp(ptr(c_full)), C.c_uint(batch),
                                                vp(_stream(stream))))

    def shard_to_full(self, world, rank, c_full, c_shard, batch, stream=None):
        check(lib().nttb200_bfv_shard_to_full(self._h, C.c_uint(world), C.c_uint(rank), vp(ptr(c_full)), vp(ptr(c_shard)), C.c_uint(batch),
                                              vp(_stream(stream))))

    def decrypt_partial_tile(self, partial, packed, c_tile, first_limb, limb_count, batch=1, stream=None):
        check(lib().nttb200_bfv_decrypt_partial_tile(self._h, vp(ptr(partial)), C.c_int(int(packed)), vp(ptr(c_tile)), C.c_uint(first_limb),
                                                     C.c_uint(limb_count), C.c_uint(batch), vp(_stream(stream))))

    def decrypt_finish_tile(self, m_out, out16, partial_sum, packed, batch=1, stream=None):
        check(lib().nttb200_bfv_decrypt_finish_tile(self._h, vp(ptr(m_out)), C.c_int(int(out16)), vp(ptr(partial_sum)), C.c_int(int(packed)),
                                                    C.c_uint(batch), vp(_stream(stream))))


def find_ntt_primes(bits: int, n: int, count: int, exclude=()):
    """C-ABI prime search (nttb200_find_ntt_primes): (primes, psi roots), largest prime first.  Host-only."""
    q, psi = (u64 * count)(), (u64 * count)()
    ex = _arr64(exclude) if exclude else None
    check(lib().nttb200_find_ntt_primes(C.c_uint(bits), C.c_uint(n), C.c_uint(count), ex, C.c_uint(len(exclude)), q, psi))
    return [int(v) for v in q], [int(v) for v in psi]


class ShardBlock(C.Structure):
    """nttb200_shard_block (include/nttb200.h)"""
    _fields_ = [("first_item", C.c_uint), ("items", C.c_uint), ("first_limb", C.c_uint), ("limb_count", C.c_uint), ("offset", C.c_size_t)]


def shard_plan(rp: int, n: int, batch: int, world: int, rank: int):
    """The calling rank's tiles: [(first_item, items, first_limb, limb_count, offset)] * world and the size of its shard buffer in words.
    Host-only (works without a GPU)."""
    blocks = (ShardBlock * world)()
    words = C.c_size_t(0)
    check(lib().nttb200_shard_plan(C.c_uint(rp), C.c_uint(n), C.c_uint(batch), C.c_uint(world), C.c_uint(rank), blocks, C.byref(words)))
    return [(b.first_item, b.items, b.first_limb, b.limb_count, b.offset) for b in blocks], int(words.value)


class Comm:
    """Communicator of the sharded BFV calls.  Comm.from_torch() adopts the NCCL communicator of the default torch.distributed
    process group (one process per GPU); Comm.single() is world size 1 (no NCCL).  A C++ host uses nttb200_comm_unique_id /
    nttb200_comm_create instead."""

    def __init__(self, handle, world, rank):
        self._h, self.world, self.rank = handle, world, rank

    @classmethod
    def single(cls):
        h = vp()
        check(lib().nttb200_comm_adopt(C.byref(h), vp(0), C.c_int(1), C.c_int(0)))
        return cls(h, 1, 0)

    @classmethod
    def fake(cls, world, rank):
        """profiling only: one rank's compute share, collectives skipped"""
        h = vp()
        check(lib().nttb200_comm_fake(C.byref(h), C.c_int(world), C.c_int(rank)))
        return cls(h, world, rank)

    @classmethod
    def from_torch(cls, group=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return cls.single()
        pg = group if group is not None else dist.distributed_c10d._get_default_group()
        backend = pg._get_backend(torch.device("cuda"))
        # the communicator is created lazily by the first collective
        t = torch.zeros(1, device="cuda")
        dist.all_reduce(t, group=group)
        torch.cuda.synchronize()
        ptr_ = int(backend._comm_ptr())
        h = vp()
        check(lib().nttb200_comm_adopt(C.byref(h), vp(ptr_), C.c_int(dist.get_world_size(group)), C.c_int(dist.get_rank(group))))
        return cls(h, dist.get_world_size(group), dist.get_rank(group))

    @classmethod
    def create(cls, world, rank, broadcast_bytes):
        """Own NCCL communicator: rank 0 draws the unique id, `broadcast_bytes(bytes_or_None) -> bytes` ships it."""
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            check(lib().nttb200_comm_unique_id(buf))
        raw = broadcast_bytes(bytes(buf) if rank == 0 else None)
        idb = (C.c_ubyte * 128).from_buffer_copy(raw)
        h = vp()
        check(lib().nttb200_comm_create(C.byref(h), idb, C.c_int(world), C.c_int(rank)))
        return cls(h, world, rank)

    def close(self):
        if self._h:
            lib().nttb200_comm_destroy(self._h)
            self._h = vp()


def keygen_rns(inp, q_amount, n, secret_key, public_key, temp, psi_table, psiinv_table, q_cons, mu_cons, q_bit_cons, stream=None):
    """bfv_keygen.cuh:95 (the unused reference parameters q, streams, mu_array, q_bit_lengths are dropped)"""
    _call("nttb200_ref_keygen_rns", vp(ptr(inp)), C.c_uint(q_amount), C.c_uint(n), vp(ptr(secret_key)), vp(ptr(public_key)), vp(ptr(temp)),
          vp(ptr(psi_table)), vp(ptr(psiinv_table)), vp(ptr(q_cons)), vp(ptr(mu_cons)), vp(ptr(q_bit_cons)), vp(_stream(stream)))


def encryption_rns(c, public_key, inp, e, n, psi_table, psiinv_table, m_poly, qi_div_t, t, q_amount, q_cons, mu_cons, q_bit_cons,
                   inv_q_last_mod_q_cons, stream=None):
    """bfv_encryption.cuh:223"""
    _call("nttb200_ref_encryption_rns", vp(ptr(c)), vp(ptr(public_key)), vp(ptr(inp)), vp(ptr(e)), C.c_uint(n), vp(ptr(psi_table)),
          vp(ptr(psiinv_table)), vp(ptr(m_poly)), vp(ptr(qi_div_t)), u64(t), C.c_uint(q_amount), vp(ptr(q_cons)), vp(ptr(mu_cons)),
          vp(ptr(q_bit_cons)), vp(ptr(inv_q_last_mod_q_cons)), vp(_stream(stream)))


def decryption_rns(c, secret_key, psi_table, psiinv_table, n, q_amount, base_change_matrix, t, gamma, mu_gamma, gamma_bits, neg_inv_t,
                   neg_inv_gamma, gamma_div_2, q_cons, mu_cons, q_bit_cons, inv_punctured_q_cons, prod_t_gamma_mod_q_cons, stream=None):
    """bfv_decryption.cuh:76 (q_amount = limbs after the drop; plaintext at c + n*(q_amount-1))"""
    _call("nttb200_ref_decryption_rns", vp(ptr(c)), vp(ptr(secret_key)), vp(ptr(psi_table)), vp(ptr(psiinv_table)), C.c_uint(n),
          C.c_uint(q_amount), vp(ptr(base_change_matrix)), u64(t), u64(gamma), u64(mu_gamma), C.c_int(gamma_bits), u64(neg_inv_t),
          u64(neg_inv_gamma), u64(gamma_div_2), vp(ptr(q_cons)), vp(ptr(mu_cons)), vp(ptr(q_bit_cons)), vp(ptr(inv_punctured_q_cons)),
          vp(ptr(prod_t_gamma_mod_q_cons)), vp(_stream(stream)))
