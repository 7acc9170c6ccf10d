"""Runs the *CUDA kernel sources* on the CPU (thread-per-CUDA-thread emulator, csrc/emu) and checks them bit-exactly
against the oracle: index maths, round schedules, twiddle indexing, the 128-byte swizzle and both arithmetic
policies are validated before any GPU time is spent.  (The PTX for TMA/mbarrier itself is only exercised on the GPU.)"""
import numpy as np
import pytest

from nttb200 import params
from tests import emu


def _ring(oracle, logn, limbs):
    n = 1 << logn
    if n in params.GET_PARAMS and limbs == 1:
        q, psi = params.GET_PARAMS[n][:2]
        qs, roots = [q], [psi]
    elif n == 32768:
        _, qs, roots = params.RNS_SETS["32k_16q"]
        qs, roots = qs[:limbs], roots[:limbs]
    elif n == 8192:
        _, qs, roots = params.RNS_SETS["8k_3q"]
        qs, roots = qs[:limbs], roots[:limbs]
    else:
        qs, roots = params.find_ntt_primes(60, n, limbs)
    tabs = [oracle.fill_psi_tables(r, q, n) for q, r in zip(qs, roots)]
    psi = np.stack([t[0] for t in tabs])
    psiinv = np.stack([t[1] for t in tabs])
    return n, qs, psi, psiinv


def _expect(oracle, a, n, qs, psi, psiinv, num, division, inverse):
    out = a.copy().reshape(num, n)
    for pidx in range(num):
        l = pidx % division
        out[pidx] = (oracle.inverse_ntt_fast(out[pidx], qs[l], psiinv[l]) if inverse
                     else oracle.forward_ntt_fast(out[pidx], qs[l], psi[l]))
    return out.reshape(-1)


CASES = [
    # logn, limbs(division), num, barrett, use_tma
    (11, 1, 2, 0, 1), (11, 1, 1, 1, 0),
    (12, 1, 2, 0, 0), (12, 1, 1, 1, 1),
    (13, 3, 4, 0, 1), (13, 3, 3, 1, 0),
    (14, 2, 2, 0, 0), (14, 1, 1, 1, 1),
    (15, 3, 4, 0, 1), (15, 2, 2, 1, 1), (15, 1, 1, 0, 0),
    (16, 2, 2, 0, 1), (16, 1, 1, 1, 0),
    (17, 1, 1, 0, 1), (17, 1, 1, 1, 0),
    # policy 2 = lazy forward (no intermediate corrections, q < 2^58); the 60-bit primes above stay on policy 0
    (11, 1, 2, 2, 1), (13, 3, 3, 2, 0), (15, 3, 3, 2, 1),
    # getParams(4096) is a 25-bit prime: floor(2^64/q) does not fit 32 bits, so the lazy forward transform takes the general final reduction
    (12, 1, 2, 2, 1),
]


@pytest.mark.parametrize("logn,limbs,num,barrett,use_tma", CASES)
def test_emu_forward_inverse(oracle, logn, limbs, num, barrett, use_tma):
    n, qs, psi, psiinv = _ring(oracle, logn, limbs)
    a = np.concatenate([oracle.fill_uniform(n, qs[pidx % limbs], 0x5EED0000 + pidx) for pidx in range(num)])
    # edge values: 0, 1, q-1 in the first polynomial
    a[0], a[1], a[2] = 0, 1, qs[0] - 1
    fwd = emu.ntt(a, n, qs, psi, psiinv, num, limbs, inverse=False, barrett=barrett, use_tma=use_tma)
    assert np.array_equal(fwd, _expect(oracle, a, n, qs, psi, psiinv, num, limbs, False))
    inv = emu.ntt(fwd, n, qs, psi, psiinv, num, limbs, inverse=True, barrett=barrett, use_tma=use_tma)
    assert np.array_equal(inv, a)


def test_emu_barrett_matches_reference_on_noncanonical_input(oracle):
    """The stateless (Barrett) path is the reference's arithmetic operation for operation, so it must agree with
    the oracle's literal restatement even for inputs >= q (e.g. the value q left behind by the `>` quirk)."""
    n, qs, psi, psiinv = _ring(oracle, 11, 1)
    q = qs[0]
    a = oracle.fill_uniform(n, q, 99)
    a[5] = q
    a[17] = q + 3
    got = emu.ntt(a, n, qs, psi, psiinv, 1, 1, inverse=False, barrett=1, use_tma=0)
    assert np.array_equal(got, oracle.forward_ntt(a, q, psi[0]))
    got = emu.ntt(a, n, qs, psi, psiinv, 1, 1, inverse=True, barrett=1, use_tma=0)
    assert np.array_equal(got, oracle.inverse_ntt(a, q, psiinv[0]))


def test_emu_barrett_reproduces_reference_barrett_glitch(oracle):
    """Found on the first GPU run: for q = 68719230977 (4k_3q limb 1, frac(2^72/q) = 0.879) the reference's single-
    correction Barrett (ntt_60bit.cuh:44-61) returns a value in [q, 2q) on this input, so the REFERENCE's INTT(NTT(a))
    != a.  The stateless path reproduces the reference bit for bit (garbage included); the Shoup path is exact."""
    n, qs, roots = params.RNS_SETS["4k_3q"]
    q, r = qs[1], roots[1]
    psi, psiinv = oracle.fill_psi_tables(r, q, n)
    a = oracle.fill_uniform(n, q, 78)
    f = oracle.forward_ntt_fast(a, q, psi)
    lit = oracle.inverse_ntt(f, q, psiinv)
    assert not np.array_equal(lit, a) and int(lit.max()) >= q          # the reference glitches here
    got = emu.ntt(f, n, [q], psi[None], psiinv[None], 1, 1, inverse=True, barrett=1, use_tma=1)
    assert np.array_equal(got, lit)
    got = emu.ntt(f, n, [q], psi[None], psiinv[None], 1, 1, inverse=True, barrett=0, use_tma=1)
    assert np.array_equal(got, a)


def test_emu_grouped_layout(oracle):
    """BFV layouts: items of [2][r][n]; transform only the second half of every item (decryption's c1) in place."""
    n, qs, psi, psiinv = _ring(oracle, 13, 3)
    r, items = 3, 2
    a = np.concatenate([oracle.fill_uniform(n, qs[(k % (2 * r)) % r], 500 + k) for k in range(items * 2 * r)])
    view = a.reshape(items, 2, r, n)
    sub = np.ascontiguousarray(a[r * n:])                      # base pointer = first polynomial of item 0, half 1
    got = emu.ntt(sub, n, qs, psi, psiinv, items * r, r, inverse=False, barrett=2, use_tma=1, group_polys=r, group_stride=2 * r * n)
    # emu.ntt copies its input, so compare against the expectation built the same way
    exp = sub.copy()
    for it in range(items):
        for l in range(r):
            off = it * 2 * r * n + l * n
            if off + n <= exp.size:
                exp[off:off + n] = oracle.forward_ntt_fast(view[it, 1, l], qs[l], psi[l])
    assert np.array_equal(got, exp)


def test_emu_lazy_policy_at_its_bound(oracle):
    """Lazy forward policy at the edge of its range: the largest 57-bit prime, n = 2^17 (17 stages, values up to 69 q),
    inputs at q - 1 everywhere (the worst case for growth) plus random ones."""
    n = 1 << 17
    qs, roots = params.find_ntt_primes(57, n, 1)
    assert qs[0].bit_length() == 57
    psi, psiinv = oracle.fill_psi_tables(roots[0], qs[0], n)
    worst = np.full(n, qs[0] - 1, dtype=np.uint64)
    rnd = oracle.fill_uniform(n, qs[0], 4242)
    a = np.concatenate([worst, rnd])
    got = emu.ntt(a, n, qs, psi[None], psiinv[None], 2, 1, inverse=False, barrett=2, use_tma=1)
    assert np.array_equal(got[:n], oracle.forward_ntt_fast(worst, qs[0], psi))
    assert np.array_equal(got[n:], oracle.forward_ntt_fast(rnd, qs[0], psi))
    # the lazy inverse (bound-tracked rounds, intermediates up to 64 q) on the same worst cases
    inv = emu.ntt(a, n, qs, psi[None], psiinv[None], 2, 1, inverse=True, barrett=2, use_tma=1)
    assert np.array_equal(inv[:n], oracle.inverse_ntt_fast(worst, qs[0], psiinv))
    assert np.array_equal(inv[n:], oracle.inverse_ntt_fast(rnd, qs[0], psiinv))
    alt = np.where(np.arange(n) % 2 == 0, qs[0] - 1, 0).astype(np.uint64)      # maximises |U - V| in the first stage
    assert np.array_equal(emu.ntt(alt, n, qs, psi[None], psiinv[None], 1, 1, inverse=True, barrett=2, use_tma=0),
                          oracle.inverse_ntt_fast(alt, qs[0], psiinv))


def test_emu_device_table_generation(oracle):
    """csrc/table_kernels.cuh (doubling construction on the device) == parameter.h:5-12 fillTablePsi128, plus companions."""
    import ctypes as C
    n, qs, roots = params.RNS_SETS["8k_3q"]
    logn, L = 13, 3
    q = np.array(qs, dtype=np.uint64)
    r = np.array(roots, dtype=np.uint64)
    ri = np.array([oracle.modinv(x, y) for x, y in zip(roots, qs)], dtype=np.uint64)
    bufs = [np.zeros(L * n, dtype=np.uint64) for _ in range(4)]
    u = C.c_ulonglong
    rc = emu.lib().emu_build_tables(*[emu.p(b, u) for b in bufs], emu.p(q, u), emu.p(r, u), emu.p(ri, u), logn, L)
    assert rc == 0
    for l in range(L):
        psi, psiinv = oracle.fill_psi_tables(roots[l], qs[l], n)
        assert np.array_equal(bufs[0][l * n:(l + 1) * n], psi) and np.array_equal(bufs[2][l * n:(l + 1) * n], psiinv)
        assert np.array_equal(bufs[1][l * n:(l + 1) * n], emu.shoup(psi, qs[l]))
        assert np.array_equal(bufs[3][l * n:(l + 1) * n], emu.shoup(psiinv, qs[l]))


@pytest.mark.parametrize("logn,limbs,num,lazy", [(11, 1, 2, 1), (12, 1, 1, 0), (13, 3, 3, 1), (15, 2, 2, 1)])
def test_emu_fused_polynomial_product(oracle, logn, limbs, num, lazy):
    """nttb200_poly_mul_batch (strided passes + ONE fused contig-forward x2 / product / contig-inverse kernel) against the
    schoolbook negacyclic product (helper.h:95-126 refPolyMul128, the check of 60bit_ntt_test.cu:85-98) at n = 2^11 and against
    the oracle's own NTT -> barrett -> INTT chain (poly_arithmetic.cuh:296-310) elsewhere; edge coefficients 0, 1, q-1."""
    n, qs, psi, psiinv = _ring(oracle, logn, limbs)
    a = np.concatenate([oracle.fill_uniform(n, qs[i % limbs], 0xA000 + i) for i in range(num)])
    b = np.concatenate([oracle.fill_uniform(n, qs[i % limbs], 0xB000 + i) for i in range(num)])
    a[0], a[1], a[2], b[0], b[1], b[2] = 0, 1, qs[0] - 1, qs[0] - 1, 0, qs[0] - 1
    got, _ = emu.polymul(a, b, n, qs, psi, psiinv, num, limbs, fwd=True, lazy=bool(lazy))
    for i in range(num):
        q, l = qs[i % limbs], i % limbs
        ai, bi = a[i * n:(i + 1) * n], b[i * n:(i + 1) * n]
        if logn == 11:
            exp = oracle.ref_poly_mul(ai, bi, q)
        else:
            exp = oracle.inverse_ntt_fast(oracle.barrett(oracle.forward_ntt_fast(ai, q, psi[l]), oracle.forward_ntt_fast(bi, q, psi[l]), q), q, psiinv[l])
        assert np.array_equal(got[i * n:(i + 1) * n], exp), f"polynomial {i}"


def test_emu_ntt_domain_product_then_inverse(oracle):
    """nttb200_ntt_domain_mul_inverse_batch: a <- INTT(a (.) b) with both operands already transformed; b untouched."""
    n, qs, psi, psiinv = _ring(oracle, 13, 3)
    num = 3
    a = np.concatenate([oracle.forward_ntt_fast(oracle.fill_uniform(n, qs[i % 3], 0xC000 + i), qs[i % 3], psi[i % 3]) for i in range(num)])
    b = np.concatenate([oracle.fill_uniform(n, qs[i % 3], 0xD000 + i) for i in range(num)])
    got, b_after = emu.polymul(a, b, n, qs, psi, psiinv, num, 3, fwd=False, lazy=True)
    assert np.array_equal(b_after, b)
    for i in range(num):
        q, l = qs[i % 3], i % 3
        exp = oracle.inverse_ntt_fast(oracle.barrett(a[i * n:(i + 1) * n], b[i * n:(i + 1) * n], q), q, psiinv[l])
        assert np.array_equal(got[i * n:(i + 1) * n], exp)


def test_emu_final_reduction_arms(oracle):
    """The lazy forward transform ends with one of three canonicalisations (ShoupLazyPolicy::fwd_final_all): q = 2^b - delta with a
    small delta (every reference prime), q > 2^32 elsewhere, and q < 2^32.  Primes around 3 * 2^(b-2) exercise the middle arm, which
    no parameter set of the reference reaches; one limb of each kind in the same batch."""
    n = 1 << 11
    q_mid, r_mid = params.find_ntt_primes(55, n, 1, below=3 << 53)
    q_pm, r_pm = params.find_ntt_primes(55, n, 1)
    q_small, r_small = params.find_ntt_primes(30, n, 1)
    qs, roots = q_mid + q_pm + q_small, r_mid + r_pm + r_small
    assert 69 * ((1 << qs[0].bit_length()) - qs[0]) > qs[0] and 69 * ((1 << 55) - qs[1]) < qs[1] and qs[2] < (1 << 32)
    tabs = [oracle.fill_psi_tables(r, q, n) for q, r in zip(qs, roots)]
    psi, psiinv = np.stack([t[0] for t in tabs]), np.stack([t[1] for t in tabs])
    num = 6
    a = np.concatenate([oracle.fill_uniform(n, qs[i % 3], 0xF1A7 + i) for i in range(num)])
    a[:3] = [0, 1, qs[0] - 1]
    fwd = emu.ntt(a, n, qs, psi, psiinv, num, 3, inverse=False, barrett=2, use_tma=1)
    assert np.array_equal(fwd, _expect(oracle, a, n, qs, psi, psiinv, num, 3, False))
    assert np.array_equal(emu.ntt(fwd, n, qs, psi, psiinv, num, 3, inverse=True, barrett=2, use_tma=1), a)


@pytest.mark.parametrize("logn,limbs,num,barrett", [(11, 3, 4, 0), (11, 2, 3, 1), (11, 3, 3, 2), (12, 3, 5, 0), (12, 1, 2, 1), (12, 3, 4, 2)])
@pytest.mark.parametrize("single", [1, 0], ids=["one_kernel", "two_kernels"])
def test_emu_small_rings_both_schedules(oracle, logn, limbs, num, barrett, single):
    """n <= 4096: the whole transform in ONE kernel (ntt_single_pass, what the library launches) and the two-kernel schedule kept for
    A/B (NTTB200_SINGLE_PASS=0) give the oracle's bits for all three arithmetic policies, several limbs, grouped polynomials."""
    if barrett == 2:                       # the lazy policies need q < 2^57
        n = 1 << logn
        qs, roots = params.find_ntt_primes(55, n, limbs)
        tabs = [oracle.fill_psi_tables(r, q, n) for q, r in zip(qs, roots)]
        psi, psiinv = np.stack([t[0] for t in tabs]), np.stack([t[1] for t in tabs])
    else:
        n, qs, psi, psiinv = _ring(oracle, logn, limbs)
    a = np.concatenate([oracle.fill_uniform(n, qs[pidx % limbs], 0xABCD00 + pidx) for pidx in range(num)])
    a[0], a[1], a[2] = 0, 1, qs[0] - 1
    emu.lib().emu_set_single_pass(single)
    try:
        fwd = emu.ntt(a, n, qs, psi, psiinv, num, limbs, inverse=False, barrett=barrett, use_tma=1)
        assert np.array_equal(fwd, _expect(oracle, a, n, qs, psi, psiinv, num, limbs, False))
        inv = emu.ntt(fwd, n, qs, psi, psiinv, num, limbs, inverse=True, barrett=barrett, use_tma=1)
        assert np.array_equal(inv, a)
    finally:
        emu.lib().emu_set_single_pass(1)
