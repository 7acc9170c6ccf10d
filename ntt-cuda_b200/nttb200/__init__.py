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

    def ntt_pass(self, a, num: int, division: int, inverse: bool, which: int, stream=None):
        """Profiling hook: only the first / second kernel of the transform (execution order)."""
        check(lib().nttb200_ntt_pass(self._h, vp(ptr(a)), C.c_uint(num), C.c_uint(division), C.c_int(int(inverse)), C.c_int(which),
                                     vp(_stream(stream))))

    # host-buffer (end-to-end) variants: numpy uint64 arrays (pinned or pageable); synchronous
    def forward_ntt_batch_host(self, a_in, a_out, num: int, division: int):
        check(lib().nttb200_forward_ntt_batch_host(self._h, vp(ptr(a_in)), vp(ptr(a_out)), C.c_uint(num), C.c_uint(division)))

    def inverse_ntt_batch_host(self, a_in, a_out, num: int, division: int):
        check(lib().nttb200_inverse_ntt_batch_host(self._h, vp(ptr(a_in)), vp(ptr(a_out)), C.c_uint(num), C.c_uint(division)))


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
