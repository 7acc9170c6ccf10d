"""CPU: the kernels of BFV ciphertext multiplication (ntt-cuda_b200/csrc/mul_kernels.cuh), compiled for the emulator, against their
definitions in Python integers.

  * k_bconv (HPS fast base conversion): bit-exact against  out_j = (sum_i y_i [B/b_i]_{o_j} - v [B]_{o_j}) mod o_j,  y_i = [x_i (B/b_i)^-1]_{b_i},
    v = floor(sum_i double(y_i) * (1/b_i) + 1/2) (the same double-precision sum, in the same order), AND semantically: the output is the
    residue vector of the centred representative of x (mod B), up to one multiple of B;
  * k_scale (HPS simple scaling) bit-exact against its definition, and = round(t/Q * d) mod p_j up to the double-precision error of its
    rounding term (the same small integer in every limb) for d given in base Q u P;
  * k_relin_accum against sum_i D_i * evk_i mod q_j.
The lazy split-word sums (Acc3) are exercised at their bound: 58-bit moduli with 15 -> 16 limbs (2 * 17 * 2^58 < 2^64)."""
import ctypes as C
import math

import numpy as np
import pytest

from nttb200 import params
from tests import emu

U64 = C.c_ulonglong


def _bases(bits, rp, k, n=2048):
    ps = params.find_ntt_primes(bits, n, rp + k)[0]
    return [int(x) for x in ps[:rp]], [int(x) for x in ps[rp:rp + k]]


def _prod(v):
    r = 1
    for x in v:
        r *= x
    return r


def _rand_res(rng, qs, n, items=1):
    return np.stack([np.stack([rng.integers(0, q, size=n, dtype=np.uint64) for q in qs]) for _ in range(items)])


def _edge(a, qs):
    """put 0, 1, q-1 into the first columns of every limb"""
    for i, q in enumerate(qs):
        a[..., i, 0] = 0
        a[..., i, 1] = 1
        a[..., i, 2] = q - 1
    return a


@pytest.mark.parametrize("bits,rp,k", [(40, 2, 3), (50, 4, 5), (58, 15, 16)])
def test_emu_bconv_matches_definition(bits, rp, k):
    n, items = 128, 2
    B_in, B_out = _bases(bits, rp, k)
    h = (bits + 1) // 2
    Bp = _prod(B_in)
    pre = [pow(Bp // b % b, -1, b) for b in B_in]
    M = [[Bp // b % o for b in B_in] for o in B_out]
    corr = [Bp % o for o in B_out]
    binv = np.array([1.0 / float(b) for b in B_in], dtype=np.float64)
    rng = np.random.default_rng(bits)
    x = _edge(_rand_res(rng, B_in, n, items), B_in)
    out = np.zeros((items, k, n), dtype=np.uint64)
    arr = lambda v: np.array(v, dtype=np.uint64)      # noqa: E731
    a_bin, a_bout, a_pre, a_M, a_corr = arr(B_in), arr(B_out), arr(pre), arr(M).reshape(-1), arr(corr)
    rc = emu.lib().emu_mul_bconv(emu.p(x, U64), emu.p(out, U64), items, rp, k, n, h, emu.p(a_bin, U64), emu.p(a_bout, U64), emu.p(a_pre, U64),
                                 emu.p(binv, C.c_double), emu.p(a_M, U64), emu.p(a_corr, U64))
    assert rc == 0
    for it in range(items):
        for j in range(0, n, 7 if n > 16 else 1):
            y = [int(x[it, i, j]) * pre[i] % B_in[i] for i in range(rp)]
            s = 0.0
            for i in range(rp):
                s = s + float(y[i]) * float(binv[i])
            v = math.floor(s + 0.5)
            for l in range(k):
                want = (sum(y[i] * M[l][i] for i in range(rp)) - v * corr[l]) % B_out[l]
                assert int(out[it, l, j]) == want, (it, j, l)
            # semantics: the centred representative of x, up to one multiple of B
            X = sum(y[i] * (Bp // B_in[i]) for i in range(rp)) % Bp
            Xc = X - Bp if X > Bp // 2 else X
            got = [int(out[it, l, j]) for l in range(k)]
            assert any(got == [(Xc + e * Bp) % o for o in B_out] for e in (0, -1, 1)), (it, j)


@pytest.mark.parametrize("bits,rp,k", [(40, 2, 3), (55, 15, 16), (58, 15, 16)])
def test_emu_scale_matches_definition(bits, rp, k):
    n, kc, t = 128, 3, 1024
    Q, P = _bases(bits, rp, k)
    h = (bits + 1) // 2
    Qp, Pp = _prod(Q), _prod(P)
    preQ = [pow((Qp // q) * Pp % q, -1, q) for q in Q]                  # (QP/q_i)^-1 mod q_i
    omega = [t * Pp // q for q in Q]
    theta = np.array([float(t * Pp % q) / float(q) for q in Q], dtype=np.float64)
    W = [[omega[i] % p for i in range(rp)] for p in P]
    lam = [t * pow(Qp, -1, p) % p for p in P]
    # d: integers below n Q^2 / 2 in magnitude, given in base Q u P (what the tensor step produces)
    rng = np.random.default_rng(1000 + bits)
    bound = n * Qp * Qp // 2
    ints = [[int(rng.integers(0, 1 << 62)) * int(rng.integers(0, 1 << 62)) % (2 * bound) - bound for _ in range(n)] for _ in range(kc)]
    ints[0][0], ints[0][1], ints[0][2] = 0, 1, -1
    d = np.array([[[v % m for v in ints[c]] for m in Q + P] for c in range(kc)], dtype=np.uint64)
    y = np.zeros((kc, k, n), dtype=np.uint64)
    arr = lambda v: np.array(v, dtype=np.uint64)      # noqa: E731
    a_Q, a_P, a_pre, a_W, a_lam = arr(Q), arr(P), arr(preQ), arr(W).reshape(-1), arr(lam)
    rc = emu.lib().emu_mul_scale(emu.p(d, U64), emu.p(y, U64), kc, rp, k, n, h, emu.p(a_Q, U64), emu.p(a_P, U64), emu.p(a_pre, U64),
                                 emu.p(theta, C.c_double), emu.p(a_W, U64), emu.p(a_lam, U64))
    assert rc == 0
    worst, tol = 0, rp * max(1, 1 << max(0, bits - 52)) + 2
    for c in range(kc):
        for j in range(0, n, 5):
            yt = [int(d[c, i, j]) * preQ[i] % Q[i] for i in range(rp)]
            f = 0.0
            for i in range(rp):
                f = f + float(yt[i]) * float(theta[i])
            v = math.floor(f + 0.5)
            for l in range(k):
                want = (sum(yt[i] * W[l][i] for i in range(rp)) + int(d[c, rp + l, j]) * lam[l] + v) % P[l]
                assert int(y[c, l, j]) == want, (c, j, l)
            # semantics: round(t / Q * d) mod p_j up to the error of the double-precision rounding term -- each product yt_i * theta_i is
            # off by at most yt_i * 2^-53 -- and THE SAME integer in every limb
            exact = (2 * t * ints[c][j] + Qp) // (2 * Qp)
            e0 = (int(y[c, 0, j]) - exact) % P[0]
            e0 = e0 - P[0] if e0 > P[0] // 2 else e0
            assert abs(e0) <= tol, (c, j, e0)
            assert all((int(y[c, l, j]) - exact - e0) % P[l] == 0 for l in range(k)), (c, j)
            worst = max(worst, abs(e0))
    assert worst <= tol


@pytest.mark.parametrize("bits,rp,items", [(40, 2, 1), (55, 5, 3), (58, 15, 2)])
def test_emu_relin_accum_matches_definition(bits, rp, items):
    n = 512
    Q, _ = _bases(bits, rp, 1)
    h = (bits + 1) // 2
    rng = np.random.default_rng(77 + bits)
    D = np.stack([np.stack([_edge(_rand_res(rng, Q, n)[0], Q) for _ in range(rp)]) for _ in range(items)])          # [items][i][j][n]
    evk = np.stack([np.stack([_rand_res(rng, Q, n)[0] for _ in range(2)]) for _ in range(rp)])                    # [i][h][j][n]
    evk[:, :, :, 1] = np.array(Q, dtype=np.uint64)[None, None, :] - 1
    acc = np.zeros((items, 2, rp, n), dtype=np.uint64)
    a_Q = np.array(Q, dtype=np.uint64)
    rc = emu.lib().emu_mul_relin_accum(emu.p(D, U64), emu.p(evk, U64), emu.p(acc, U64), n, rp, items, h, emu.p(a_Q, U64))
    assert rc == 0
    for it in range(items):
        for hh in range(2):
            for j in range(rp):
                want = np.zeros(n, dtype=object)
                for i in range(rp):
                    want = (want + D[it, i, j].astype(object) * evk[i, hh, j].astype(object)) % Q[j]
                assert np.array_equal(acc[it, hh, j].astype(object), want), (it, hh, j)
