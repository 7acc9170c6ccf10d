"""oracle/bfv_mul_oracle.py -- TEST INFRASTRUCTURE: exact big-integer BFV ciphertext multiplication and relinearisation.

The reference (ozgunozerk/NTT-Cuda) stops at decryption; homomorphic multiplication is the paper's stated future work (Article.pdf
p.29) and SURVEY.md 8f-4's last row.  There is therefore no reference code to restate: this oracle is the TEXTBOOK definition of the
operation (Fan-Vercauteren 2012, section 4) computed with unbounded Python integers --

    tensor   d0 = a0 b0, d1 = a0 b1 + a1 b0, d2 = a1 b1   over Z[X]/(X^n + 1), on the centred lifts of the ciphertext components
    scaling  y_k = round(t / Q * d_k)  (round half up), then reduced mod every q_i
    relin    c_h = y_h + sum_i [y2]_{q_i} * evk_{i,h}     (RNS digits; evk_i = (-(a_i s + e_i) + g_i s^2, a_i))
    decrypt  m = round(t / Q * [y0 + y1 s + y2 s^2]_Q) mod t        (degree-2 decryption, for checking the tensor step alone)

-- so it is independent of the GPU's RNS / floating-point route (Halevi-Polyakov-Shoup base extension + simple scaling,
ntt-cuda_b200/csrc/mul_kernels.cuh).  The GPU's scaled tensor may differ from the exact one by a small integer per coefficient (its
rounding term is a double-precision sum; bound in tests/test_gpu_mul.py); everything after it is integer arithmetic and must match
bit for bit.  Only tests/ import this module.  Parity status: no reference vector exists for this operation ("parity unpinned" in the
sense of the task statement); it is pinned instead to the definition and to Dec(c_a * c_b) = m_a * m_b.
"""
import numpy as np


def _centered(x, m):
    x %= m
    return x - m if x > m // 2 else x


def crt_lift(res, qs):
    """res[len(qs)][n] residues -> list of n centred Python integers modulo Q = prod(qs)."""
    Q = 1
    for q in qs:
        Q *= int(q)
    out = [0] * len(res[0])
    for i, q in enumerate(qs):
        q = int(q)
        Qi = Q // q
        w = Qi * pow(Qi % q, -1, q)
        col = [int(v) for v in res[i]]
        out = [o + w * v for o, v in zip(out, col)]
    return [_centered(v, Q) for v in out], Q


def negacyclic_mul(a, b):
    """Product of two integer polynomials (lists of Python ints, signed) modulo X^n + 1, by Kronecker substitution."""
    n = len(a)
    bound = max(1, max(abs(v) for v in a)) * max(1, max(abs(v) for v in b)) * n
    B = bound.bit_length() + 2
    pa = sum(v << (B * i) for i, v in enumerate(a))
    pb = sum(v << (B * i) for i, v in enumerate(b))
    prod = pa * pb
    half, mod = 1 << (B - 1), 1 << B
    c = [0] * (2 * n)
    for k in range(2 * n):
        d = prod & (mod - 1)
        if d >= half:
            d -= mod
        c[k] = d
        prod = (prod - d) >> B
    return [c[k] - c[k + n] for k in range(n)]


def round_div(x, q):
    """round(x / q), halves up (floor(x / q + 1/2))."""
    return (2 * x + q) // (2 * q)


def tensor_scaled(ca, cb, n, qs, t):
    """ca, cb: numpy uint64 [2][len(qs)][n] (coefficient domain, canonical).  Returns (y: list of 3 lists of n centred-free Python ints
    = round(t / Q * d_k), Q)."""
    a, Q = zip(*[crt_lift(ca[h], qs) for h in range(2)])
    b, _ = zip(*[crt_lift(cb[h], qs) for h in range(2)])
    Q = Q[0]
    d0 = negacyclic_mul(a[0], b[0])
    d1 = [x + y for x, y in zip(negacyclic_mul(a[0], b[1]), negacyclic_mul(a[1], b[0]))]
    d2 = negacyclic_mul(a[1], b[1])
    return [[round_div(t * v, Q) for v in d] for d in (d0, d1, d2)], Q


def to_rns(y, qs):
    """list of n Python ints -> numpy uint64 [len(qs)][n]"""
    return np.array([[v % int(q) for v in y] for q in qs], dtype=np.uint64)


def decrypt_degree2(y, s_coeff, Q, t):
    """y: 3 lists of Python ints (any representatives mod Q); s_coeff: secret key as signed Python ints.  Exact BFV decryption."""
    s2 = negacyclic_mul(s_coeff, s_coeff)
    acc = [a + b + c for a, b, c in zip(y[0], negacyclic_mul(y[1], s_coeff), negacyclic_mul(y[2], s2))]
    return np.array([round_div(t * _centered(v, Q), Q) % t for v in acc], dtype=np.uint64)


def secret_key_coefficients(orc, R, sk):
    """sk[r][n] in the NTT domain (as keygen leaves it) -> signed ternary-ish coefficients (from limb 0)."""
    q0 = int(R.q[0])
    c = orc.inverse_ntt_fast(np.ascontiguousarray(sk[:R.n]), q0, R.psiinv[0])
    return [_centered(int(v), q0) for v in c]


def plain_product(ma, mb, t):
    """m_a * m_b mod (X^n + 1, t)"""
    p = negacyclic_mul([int(v) for v in ma], [int(v) for v in mb])
    return np.array([v % t for v in p], dtype=np.uint64)


def relinearize(orc, R, y_rns, evk):
    """y_rns: numpy uint64 [3][rp][n] (coefficient domain); evk: numpy uint64 [rp][2][rp][n] (NTT domain).  Returns c[2][rp][n]:
    c_h = y_h + sum_i INTT( NTT([y2]_{q_i} mod q_j) (.) evk[i][h][j] )  -- exact modular arithmetic limb by limb."""
    rp, n = R.r - 1, R.n
    out = np.zeros((2, rp, n), dtype=np.uint64)
    for j in range(rp):
        qj = int(R.q[j])
        acc = [np.zeros(n, dtype=object), np.zeros(n, dtype=object)]
        for i in range(rp):
            digit = (y_rns[2][i].astype(object) % qj).astype(np.uint64)
            dh = orc.forward_ntt_fast(np.ascontiguousarray(digit), qj, R.psi[j]).astype(object)
            for h in range(2):
                acc[h] = (acc[h] + dh * evk[i][h][j].astype(object)) % qj
        for h in range(2):
            back = orc.inverse_ntt_fast(np.ascontiguousarray(acc[h].astype(np.uint64)), qj, R.psiinv[j]).astype(object)
            out[h][j] = ((back + y_rns[h][j].astype(object)) % qj).astype(np.uint64)
    return out


def relin_keygen(orc, R, sk, seed=1):
    """A relinearisation key with numpy randomness (CPU-only tests; the GPU draws its own from Salsa20): evk[rp][2][rp][n], NTT domain."""
    rp, n = R.r - 1, R.n
    rng = np.random.default_rng(seed)
    evk = np.zeros((rp, 2, rp, n), dtype=np.uint64)
    for i in range(rp):
        e = rng.integers(-8, 9, size=n)
        for j in range(rp):
            qj = int(R.q[j])
            a = rng.integers(0, qj, size=n, dtype=np.uint64)
            s = sk[j * n:(j + 1) * n].astype(object)
            eh = orc.forward_ntt_fast(np.ascontiguousarray((e % qj).astype(np.uint64)), qj, R.psi[j]).astype(object)
            v = (-(a.astype(object) * s + eh)) % qj
            if i == j:
                v = (v + s * s) % qj
            evk[i][0][j] = v.astype(np.uint64)
            evk[i][1][j] = a
    return evk
