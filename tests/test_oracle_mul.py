"""CPU: the exact big-integer BFV multiplication / relinearisation oracle (oracle/bfv_mul_oracle.py) is self-consistent with the
reference-restating oracle's encryption and decryption: Dec2(tensor(c_a, c_b)) = m_a m_b and Dec(relin(tensor)) = m_a m_b."""
import numpy as np

from nttb200 import params
from oracle import bfv_mul_oracle as mo


def _ring(oracle, n=2048, bits=40, limbs=4):
    qs, roots = params.find_ntt_primes(bits, n, limbs)
    return oracle.Ring(n, qs, roots)


def test_kronecker_negacyclic_product_small():
    a, b = [1, -2, 3, 4], [5, 6, -7, 8]
    n = 4
    ref = [0] * n
    for i in range(n):
        for j in range(n):
            k = i + j
            ref[k % n] += a[i] * b[j] * (-1 if k >= n else 1)
    assert mo.negacyclic_mul(a, b) == ref


def test_exact_multiply_and_relinearize_decrypt_to_the_product(oracle):
    R = _ring(oracle)
    n, r = R.n, R.r
    rp = r - 1
    sk, pk, _, _ = oracle.keygen_rns(R)
    ma, mb = oracle.fill_uniform(n, R.t, 0xA1), oracle.fill_uniform(n, R.t, 0xB2)
    oracle.set_nonce(1)
    ca, _ = oracle.encryption_rns(R, pk, ma)
    oracle.set_nonce(2)
    cb, _ = oracle.encryption_rns(R, pk, mb)
    oracle.set_nonce(0)
    qs = [int(x) for x in R.q[:rp]]
    A = ca.reshape(2, r, n)[:, :rp]
    B = cb.reshape(2, r, n)[:, :rp]
    y, Q = mo.tensor_scaled(A, B, n, qs, R.t)
    expect = mo.plain_product(ma, mb, R.t)
    s = mo.secret_key_coefficients(oracle, R, sk)
    assert set(s) <= {-1, 0, 1, 2}
    assert np.array_equal(mo.decrypt_degree2(y, s, Q, R.t), expect)
    # relinearise with an oracle-made key, decrypt with the reference-restating decryption
    evk = mo.relin_keygen(oracle, R, sk)
    y_rns = np.stack([mo.to_rns(v, qs) for v in y])
    c2 = mo.relinearize(oracle, R, y_rns, evk)
    full = np.zeros((2, r, n), dtype=np.uint64)
    full[:, :rp] = c2
    plain, _ = oracle.decryption_rns(R, full.reshape(-1), sk)
    assert np.array_equal(plain, expect)


def test_find_ntt_primes_c_entry_on_the_host():
    """nttb200_find_ntt_primes is host-only: the library's search equals the Python search used since round 1, roots are primitive."""
    import nttb200
    for bits, n, cnt in ((55, 65536, 3), (40, 2048, 4)):
        q, psi = nttb200.find_ntt_primes(bits, n, cnt)
        assert (q, psi) == params.find_ntt_primes(bits, n, cnt)
        for qq, pp in zip(q, psi):
            assert qq % (2 * n) == 1 and pow(pp, n, qq) == qq - 1
    q, _ = nttb200.find_ntt_primes(55, 32768, 2, exclude=[36028797017456641])
    assert 36028797017456641 not in q
