"""GPU parity of the entry points round 1 only exercised on the CPU emulator ("What's weak" 1-2):
nttb200_barrett_batch_3param, nttb200_convert_ternary_gaussian_x2, nttb200_poly_add_negate_xq,
nttb200_divide_and_round_q_last_inplace_loop, nttb200_gaussian_dist, and the 58-bit C1 prime at n = 4096 in the pytest suite."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from nttb200 import params  # noqa: E402


def test_barrett_batch_3param_and_add_negate(oracle):
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = params.RNS_SETS["4k_3q"]
    R = oracle.Ring(n, qs, roots)
    qd, mud, qbd = to_dev(R.qa), to_dev(R.mu), to_dev(R.qbit)
    A = np.concatenate([oracle.fill_uniform(n, int(R.q[p % 3]), 10 + p) for p in range(6)])
    B = np.concatenate([oracle.fill_uniform(n, int(R.q[p % 3]), 20 + p) for p in range(6)])
    A[0], B[0] = R.q[0] - 1, R.q[0] - 1
    c = torch.zeros(6 * n, dtype=torch.int64, device="cuda")
    nttb200.barrett_batch_3param(c, to_dev(A), to_dev(B), n, 6, 3, qd, mud, qbd)
    assert np.array_equal(to_host(c), oracle.barrett_batch_3param(A, B, n, 6, 3, R.qa, R.mu, R.qbit))
    a3 = to_dev(A[:3 * n])
    nttb200.poly_add_negate_xq(a3, to_dev(B[:3 * n]), n, 3, qd)
    assert np.array_equal(to_host(a3), oracle.poly_add_negate_xq(A[:3 * n], B[:3 * n], n, qs))


def test_divide_and_round_q_last_inplace_loop(oracle):
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = params.RNS_SETS["8k_4q"]
    R = oracle.Ring(n, qs, roots)
    q0, ql = int(R.q[0]), int(R.q[-1])
    a = oracle.fill_uniform(n, q0, 3)
    last = oracle.fill_uniform(n, ql, 4)
    a[:2] = [0, q0 - 1]
    last[:2] = [ql - 1, 0]
    half_mod = (ql >> 1) % q0
    d = to_dev(a)
    nttb200.divide_and_round_q_last_inplace_loop(d, to_dev(last), n, q0, half_mod, int(R.inv_q_last_mod_q[0]), int(R.mu[0]), int(R.qbit[0]))
    assert np.array_equal(to_host(d), oracle.divide_and_round_q_last_inplace_loop(a, last, q0, half_mod, int(R.inv_q_last_mod_q[0]),
                                                                                   int(R.mu[0]), int(R.qbit[0])))


def test_gaussian_dist_and_convert_ternary_gaussian_x2(oracle):
    """The two samplers that involve normcdfinvf: values must equal the CPU inverse-normal except (rarely) at a truncation boundary, and the
    two GPU entry points must agree with EACH OTHER exactly (same device function)."""
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = params.RNS_SETS["4k_3q"]
    r = len(qs)
    qd = to_dev(np.array(qs, dtype=np.uint64))
    inb = oracle.generate_random_default(9 * n)
    ind = to_dev(inb)
    c = torch.zeros(2 * r * n, dtype=torch.int64, device="cuda")
    e = torch.zeros(2 * r * n, dtype=torch.int64, device="cuda")
    nttb200.convert_ternary_gaussian_x2(ind, c, e, n, r, qd)
    hc, he = to_host(c).reshape(2, r, n), to_host(e).reshape(2, r, n)
    tern = oracle.ternary_dist_xq(inb, n, qs).reshape(r, n)
    assert np.array_equal(hc[0], tern) and np.array_equal(hc[1], tern)                  # u copied into both halves (bfv_encryption.cuh:23-36)
    for h, off in ((0, n), (1, 5 * n)):
        w = inb[off:off + 4 * n].view(np.uint32)
        g = torch.zeros(n, dtype=torch.int64, device="cuda")
        for l in range(r):
            nttb200.gaussian_dist(to_dev(w), g, n, None, qs[l])
            assert np.array_equal(to_host(g), he[h, l])                                 # GPU == GPU, exact
            cpu = oracle.convert_gaussian(w, qs[l])
            assert np.count_nonzero(cpu != he[h, l]) <= 3


def test_c1_58bit_prime_round_trip_and_oracle(oracle):
    """BASELINE config 1 with the prime it means: N = 4096, q = 288230376135196673 (parameter.h:43-47), through the context path
    (general Shoup policy: 58 bits is above the lazy bound) and the stateless reference-contract path, against the oracle."""
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n = 4096
    q, psi = params.GET_PARAMS_4096_58BIT[:2]
    tab, tabinv = oracle.fill_psi_tables(psi, q, n)
    a = oracle.fill_uniform(n, q, 0xC1)
    a[:2] = [0, q - 1]
    exp = oracle.forward_ntt(a, q, tab)
    ctx = nttb200.Context(n, [q], [psi])
    d = to_dev(a)
    ctx.forward_ntt_batch(d, 1, 1)
    assert np.array_equal(to_host(d), exp)
    ctx.inverse_ntt_batch(d, 1, 1)
    assert np.array_equal(to_host(d), a)
    ctx.close()
    d2 = to_dev(a)
    nttb200.forwardNTT(d2, n, None, q, oracle.mu(q), oracle.qbit(q), to_dev(tab))
    assert np.array_equal(to_host(d2), exp)
    nttb200.inverseNTT(d2, n, None, q, oracle.mu(q), oracle.qbit(q), to_dev(tabinv))
    assert np.array_equal(to_host(d2), a)
