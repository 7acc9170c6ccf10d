"""GPU: BFV ciphertext x ciphertext multiplication + relinearisation through the C ABI.
  * semantic parity for 4k_3q / 8k_4q / 32k_16q:  Dec(relin(c_a * c_b)) = m_a * m_b mod (X^n + 1, t), batch, squaring, aliasing
  * against the exact big-integer oracle (oracle/bfv_mul_oracle.py) for 4k_3q / 8k_4q: the scaled tensor equals round(t/Q d) up to a
    small integer that is THE SAME in every limb (the double-precision rounding term of the RNS scaling), and relinearisation --
    integer arithmetic only -- is bit-exact given the same degree-2 ciphertext and key
  * nttb200_find_ntt_primes against the Python search used since round 1."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from nttb200 import params  # noqa: E402
from oracle import bfv_mul_oracle as mo  # noqa: E402


def _setup(oracle, name, B):
    import torch
    import nttb200
    from tests.gpu_util import to_dev
    n, q, roots = params.RNS_SETS[name]
    R = oracle.Ring(n, q, roots)
    bfv = nttb200.Bfv(n, q, roots)
    rn = R.r * n
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk)
    bfv.load_keys(sk, pk)
    bfv.relin_keygen(sk)
    ma = np.concatenate([oracle.fill_uniform(n, R.t, 0x111 + k) for k in range(B)])
    mb = np.concatenate([oracle.fill_uniform(n, R.t, 0x222 + k) for k in range(B)])
    ca = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    cb = torch.zeros_like(ca)
    bfv.encrypt(ca, pk, to_dev(ma), batch=B, nonce0=100)
    bfv.encrypt(cb, pk, to_dev(mb), batch=B, nonce0=200)
    return R, bfv, sk, ma, mb, ca, cb


@pytest.mark.parametrize("name", ["4k_3q", "8k_4q", "16k_9q", "32k_16q"])
def test_multiply_relinearize_decrypts_to_the_product(oracle, name):
    import torch
    from tests.gpu_util import to_host
    B = 2
    R, bfv, sk, ma, mb, ca, cb = _setup(oracle, name, B)
    n, rn = R.n, R.r * R.n
    prod = torch.full((B * 2 * rn,), -1, dtype=torch.int64, device="cuda")
    bfv.mul(prod, ca, cb, batch=B)
    out = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    bfv.decrypt(out, prod.clone(), None, batch=B)
    got = to_host(out)
    for k in range(B):
        assert np.array_equal(got[k * n:(k + 1) * n], mo.plain_product(ma[k * n:(k + 1) * n], mb[k * n:(k + 1) * n], R.t)), (name, k)
    # the padding limb of the output is left alone
    assert (prod.view(B, 2, R.r, n)[:, :, R.r - 1] == -1).all()
    # squaring (same pointer twice) and in-place output
    sq = ca.clone()
    bfv.mul(sq, sq, sq, batch=B)
    bfv.decrypt(out, sq, None, batch=B)
    got = to_host(out)
    assert np.array_equal(got[:n], mo.plain_product(ma[:n], ma[:n], R.t))
    # a product is a ciphertext like any other: add the other operand and decrypt  (m_a m_b + m_b)
    s = prod.clone()
    s.view(B, 2, R.r, n)[:, :, R.r - 1] = 0
    bfv.add(s, cb, batch=B)
    bfv.decrypt(out, s, None, batch=B)
    exp = (mo.plain_product(ma[:n], mb[:n], R.t) + mb[:n]) % np.uint64(R.t)
    assert np.array_equal(to_host(out)[:n], exp)
    bfv.close()


@pytest.mark.parametrize("name", ["4k_3q", "8k_4q"])
def test_tensor_and_relinearization_against_the_exact_oracle(oracle, name):
    import torch
    import nttb200
    from tests.gpu_util import to_host
    R, bfv, sk, ma, mb, ca, cb = _setup(oracle, name, 1)
    n, r = R.n, R.r
    rp = r - 1
    qs = [int(x) for x in R.q[:rp]]
    y = torch.zeros(3 * rp * n, dtype=torch.int64, device="cuda")
    bfv.mul_tensor(y, ca, cb, batch=1)
    gy = to_host(y).reshape(3, rp, n)
    A = to_host(ca).reshape(2, r, n)[:, :rp]
    Bm = to_host(cb).reshape(2, r, n)[:, :rp]
    ey, Q = mo.tensor_scaled(A, Bm, n, qs, R.t)
    worst = 0
    for c in range(3):
        ex = mo.to_rns(ey[c], qs)
        d0 = None
        for i, q in enumerate(qs):
            d = (gy[c][i].astype(object) - ex[i].astype(object)) % q
            d = np.array([int(v) - q if int(v) > q // 2 else int(v) for v in d], dtype=np.int64)
            if d0 is None:
                d0 = d
            assert np.array_equal(d, d0), "the GPU's scaled tensor is not one integer vector: limbs disagree"
        worst = max(worst, int(np.abs(d0).max()))
    assert worst <= 2 * rp + 2, worst          # |sum_i yt_i theta_i - its double-precision value| stays below rp * 2^55 * 2^-52 + 1/2
    # degree-2 decryption of the GPU's tensor, exactly
    s = mo.secret_key_coefficients(oracle, R, to_host(sk))
    ylift = [mo.crt_lift(gy[c], qs)[0] for c in range(3)]
    assert np.array_equal(mo.decrypt_degree2(ylift, s, Q, R.t), mo.plain_product(ma, mb, R.t))
    # relinearisation is integer-only: bit-exact against the oracle on the same inputs and key
    addr, words = bfv.relin_key()
    evk = nttb200.download(addr, words).reshape(rp, 2, rp, n)
    out = torch.zeros(2 * r * n, dtype=torch.int64, device="cuda")
    bfv.relinearize(out, y, batch=1)
    exp = mo.relinearize(oracle, R, gy, evk)
    assert np.array_equal(to_host(out).reshape(2, r, n)[:, :rp], exp)
    bfv.close()


def test_find_ntt_primes_c_entry_matches_python_search():
    import nttb200
    for bits, n, cnt in ((55, 65536, 4), (40, 2048, 3), (57, 131072, 2)):
        assert nttb200.find_ntt_primes(bits, n, cnt) == params.find_ntt_primes(bits, n, cnt)
    q, _ = nttb200.find_ntt_primes(55, 32768, 2, exclude=[36028797017456641])
    assert 36028797017456641 not in q
