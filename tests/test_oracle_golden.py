"""Pins the CPU oracle against the reference's own fixtures and identities.

(i)   decryption_test.cu KAT: c_host/sk_host -> plaintext i % 10            (the only golden vector)
(ii)  NTT -> pointwise -> INTT == schoolbook negacyclic product            (60bit_ntt_test.cu:85-98, check=1)
(iii) keygen -> encrypt -> decrypt round trip identity                       (demo.cu:302-311)
(iv)  Barrett vs plain `%`                                                   (old/barrett_demo.cu)
(v)   Salsa20/20 against the published specification's test vector.
"""
import os

import numpy as np
import pytest

from nttb200 import params

GOLD = os.path.join(os.path.dirname(__file__), "golden", "decryption_kat.npz")


def test_decryption_kat(oracle):
    g = np.load(GOLD)
    ring = oracle.Ring(int(g["n"]), [int(x) for x in g["q"]], [int(x) for x in g["psi_roots"]], t=int(g["t"]), gamma=int(g["gamma"]))
    plain, _ = oracle.decryption_rns(ring, g["c_host"], g["sk_host"])
    expect = np.arange(4096, dtype=np.uint64) % 10
    assert np.array_equal(plain, expect)


def test_kat_padding_is_ignored(oracle):
    """decryption_test.cu:349-354: c_host[8192:12288] and [20480:24576] are padding."""
    g = np.load(GOLD)
    ring = oracle.Ring(4096, [int(x) for x in g["q"]], [int(x) for x in g["psi_roots"]])
    c = g["c_host"].copy()
    c[8192:12288] = 0
    c[20480:24576] = 12345
    plain, _ = oracle.decryption_rns(ring, c, g["sk_host"])
    assert np.array_equal(plain, np.arange(4096, dtype=np.uint64) % 10)


@pytest.mark.parametrize("n", [2048, 4096])
def test_polymul_vs_schoolbook(oracle, n):
    q, psi, psiinv, ninv, qb = params.GET_PARAMS[n]
    assert oracle.qbit(q) == qb and oracle.modinv(psi, q) == psiinv and oracle.modinv(n, q) == ninv
    tab, tabinv = oracle.fill_psi_tables(psi, q, n)
    a = oracle.fill_uniform(n, q, 1)
    b = oracle.fill_uniform(n, q, 2)
    prod = oracle.inverse_ntt(oracle.barrett(oracle.forward_ntt(a, q, tab), oracle.forward_ntt(b, q, tab), q), q, tabinv)
    assert np.array_equal(prod, oracle.ref_poly_mul(a, b, q))


def test_58bit_prime_roundtrip_and_fast_path(oracle):
    q, psi, psiinv, ninv, qb = params.GET_PARAMS_4096_58BIT
    n = 4096
    assert oracle.qbit(q) == qb and oracle.modinv(psi, q) == psiinv
    tab, tabinv = oracle.fill_psi_tables(psi, q, n)
    a = oracle.fill_uniform(n, q, 7)
    f = oracle.forward_ntt(a, q, tab)
    assert np.array_equal(f, oracle.forward_ntt_fast(a, q, tab))
    assert np.array_equal(oracle.inverse_ntt(f, q, tabinv), a)
    assert np.array_equal(oracle.inverse_ntt_fast(f, q, tabinv), a)
    # definition check on a few outputs: out[bitrev(k)] = sum_j a_j psi^((2k+1) j)
    for k in (0, 1, 77, n - 1):
        w = pow(psi, 2 * k + 1, q)
        acc, p = 0, 1
        for j in range(n):
            acc = (acc + int(a[j]) * p) % q
            p = p * w % q
        assert int(f[int(oracle.lib().orc_bitrev(k, 12))]) == acc


def test_psi_table_definition(oracle):
    q, psi = params.GET_PARAMS[2048][:2]
    tab, tabinv = oracle.fill_psi_tables(psi, q, 2048)
    psiinv = pow(psi, q - 2, q)
    for i in (0, 1, 2, 3, 1000, 2047):
        br = int(oracle.lib().orc_bitrev(i, 11))
        assert int(tab[i]) == pow(psi, br, q) and int(tabinv[i]) == pow(psiinv, br, q)
    assert pow(psi, 2048, q) == q - 1


def test_barrett_matches_mod(oracle):
    rng = np.random.default_rng(0)
    for q in (33538049, 137438691329, 36028797017456641, 288230376135196673, params.GAMMA, (1 << 61) - 1 + 2):
        qb = oracle.qbit(q)
        m = oracle.mu(q, qb)
        xs = [0, 1, q - 1, q // 2] + [int(v) % q for v in rng.integers(0, 2**63, 200)]
        for a in xs:
            for b in (0, 1, q - 1, xs[7], xs[11]):
                assert int(oracle.lib().orc_barrett_mul(a, b, q, m, qb)) == a * b % q


def test_derived_params_demo_comments(oracle):
    """demo.cu keeps a few hand-computed constants for the (4k,3q) set in comments (:83, :90, :116)."""
    n, q, roots = params.RNS_SETS["4k_3q"]
    ring = oracle.Ring(n, q, roots)
    assert [int(x) for x in ring.qi_div_t] == [67108792, 67108624, 134217600]
    assert [int(x) for x in ring.inv_punctured_q] == [26179219651, 42540076863]
    assert [int(x) for x in ring.prod_t_gamma_mod_q] == [37067052033, 64547873793]
    for i in range(3):
        assert pow(int(roots[i]), n, q[i]) == q[i] - 1


def test_salsa20_spec_vector(oracle):
    """Salsa20/20 256-bit key, ECRYPT verified test vector set 1, vector 0: key = 80 00.., IV = 0."""
    key = bytes([0x80] + [0] * 31)
    ks = oracle.salsa20_keystream(64, key, 0)
    assert bytes(ks).hex().upper() == ("E3BE8FDD8BECA2E3EA8EF9475B29A6E7003951E1097A5C38D23B7A5FAD9F6844"
                                       "B22C97559E2723C7CBBD3FE4FC8D9A0744652A83E72A9C461876AF4D7EF1A117")


def test_ternary_and_uniform_semantics(oracle):
    q = 274877562881
    inb = np.arange(256, dtype=np.uint8)
    tv = oracle.ternary_dist_xq(inb, 256, [q])
    assert int(tv[0]) == q - 1 and int(tv[84]) == q - 1 and int(tv[85]) == 0 and int(tv[169]) == 0
    assert int(tv[170]) == 1 and int(tv[254]) == 1 and int(tv[255]) == 2          # byte 255 -> 2 (reference quirk)
    legacy = oracle.convert_ternary(inb, q)
    assert int(legacy[85]) == q - 1 and int(legacy[86]) == 0 and int(legacy[171]) == 1 and int(legacy[255]) == 1
    u = oracle.convert_range(np.array([0, 1 << 63, (1 << 64) - 1], dtype=np.uint64), q)
    assert int(u[0]) == 0 and int(u[1]) == (q - 1) // 2 and int(u[2]) == q - 1
    g = oracle.gaussian_samples(np.array([0, 1 << 31, (1 << 32) - 1, 1 << 30], dtype=np.uint32))
    assert g[1] == 0 and g[0] <= -16 and g[2] >= 16 and g[3] == -2


@pytest.mark.parametrize("name", ["4k_3q", "8k_3q"])
def test_keygen_encrypt_decrypt_roundtrip(oracle, name):
    n, q, roots = params.RNS_SETS[name]
    ring = oracle.Ring(n, q, roots)
    sk, pk, temp, inb = oracle.keygen_rns(ring)
    m = oracle.fill_uniform(n, ring.t, 0xC0FFEE)
    c, e = oracle.encryption_rns(ring, pk, m)
    plain, _ = oracle.decryption_rns(ring, c, sk)
    assert np.array_equal(plain, m)
    # keystream is identical in keygen and encryption: u == s before the NTT
    r = ring.r
    s_nat = oracle.inverse_ntt(sk[:n], q[0], ring.psiinv[0])
    assert np.array_equal(s_nat, oracle.ternary_dist_xq(inb[:n], n, [q[0]]))
    # public key relation: pk0 + a*s + e == 0 (NTT domain, limb 0)
    a_s = oracle.barrett(pk[r * n: r * n + n], sk[:n], q[0])
    lhs = oracle.inverse_ntt((pk[:n] + a_s) % np.uint64(q[0]), q[0], ring.psiinv[0])
    assert np.array_equal((lhs + temp[:n]) % np.uint64(q[0]), np.zeros(n, dtype=np.uint64))


def test_reference_barrett_exactness_criterion(oracle):
    """the reference's one-correction Barrett is exact for all a, b < q when frac(2^(2 qbit)/q) < 1/2 (the fused
    kernels rely on this to replace it by Shoup products) and can be 1*q off otherwise (a witness is kept for 68719230977)."""
    rng = np.random.default_rng(5)
    for name in ("32k_16q", "16k_5q", "8k_4q", "8k_3q", "4k_3q"):
        for q in params.RNS_SETS[name][1]:
            qb, m = oracle.qbit(q), oracle.mu(q)
            delta = ((1 << (2 * qb)) % q) / q
            if delta >= 0.5:
                continue
            # adversarial operands: products just above a multiple of q (small true remainder) and near 2^(2 qbit)
            xs = [q - 1, q - 2, (q - 1) // 2, 1 << (qb - 1)] + [int(v) % q for v in rng.integers(1, 2**63, 300)]
            for a in xs[:40]:
                inv = pow(a, q - 2, q)
                for r in (0, 1, 2, 3, q - 1):
                    b = inv * r % q                      # a * b = r (mod q)
                    assert int(oracle.lib().orc_barrett_mul(a, b, q, m, qb)) == r
            for a, b in zip(xs[4:], xs[5:]):
                assert int(oracle.lib().orc_barrett_mul(a, b, q, m, qb)) == a * b % q
    # a witness of the glitch for the delta = 0.879 prime: result in [q, 2q)
    q = 68719230977
    qb, m = oracle.qbit(q), oracle.mu(q)
    assert ((1 << (2 * qb)) % q) / q > 0.75
    found = False
    for a in range(q - 1, q - 4000, -1):
        b = q - 1
        if int(oracle.lib().orc_barrett_mul(a, b, q, m, qb)) != a * b % q:
            found = True
            break
    assert found


# ---- (vi) outputs of the reference ITSELF, run on a B200 (scripts/make_reference_fixtures.py -> tests/golden/reference_gpu_fixtures*) ----
def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _fixtures():
    import json
    d = os.path.join(os.path.dirname(__file__), "golden")
    return json.load(open(os.path.join(d, "reference_gpu_fixtures.json"))), np.load(os.path.join(d, "reference_gpu_fixtures_draws.npz"))


@pytest.mark.parametrize("name", ["4k_3q", "8k_3q", "32k_16q"])
def test_oracle_ntt_matches_reference_gpu_outputs(oracle, name):
    """Raw forward-NTT output values are pinned by no test of the reference; here they are pinned against the
    reference's own kernels (forwardNTT_batch / inverseNTT_batch rebuilt for sm_100a, run on a B200): SHA-256 per polynomial."""
    fx, _ = _fixtures()
    rec = fx["ntt"][name]
    n, qs, roots = params.RNS_SETS[name]
    r, num = len(qs), rec["num"]
    tabs = [oracle.fill_psi_tables(roots[l], qs[l], n) for l in range(r)]
    fwd = []
    for p in range(num):
        a = oracle.fill_uniform(n, qs[p % r], 0x5EED0000 + p)
        f = oracle.forward_ntt_fast(a, qs[p % r], tabs[p % r][0])
        assert _sha(f) == rec["fwd_sha256_per_poly"][p], f"polynomial {p}"
        assert np.array_equal(oracle.inverse_ntt_fast(f, qs[p % r], tabs[p % r][1]), a)
        fwd.append(f)
    fwd = np.concatenate(fwd)
    assert [int(v) for v in fwd[:8]] == rec["fwd_head"] and _sha(fwd) == rec["fwd_sha256"]


@pytest.mark.parametrize("name", ["4k_3q", "8k_4q", "16k_5q"])
def test_oracle_bfv_matches_reference_gpu_outputs(oracle, name):
    """keygen_rns -> encryption_rns -> decryption_rns of the reference on a B200: keystream, secret key, public key, ciphertext
    (padding limb included) and plaintext digests.  The gaussian draws are taken from the fixture (normcdfinvf is the one value
    a CPU cannot reproduce bit for bit; the oracle's own draws may differ from them at a few truncation boundaries only)."""
    fx, draws = _fixtures()
    rec = fx["bfv"][name]
    n, qs, roots = params.RNS_SETS[name]
    R = oracle.Ring(n, qs, roots)
    e = draws[f"{name}_keygen_e"].astype(np.int32)
    sk, pk, temp, inb = oracle.keygen_rns(R, e_samples=np.ascontiguousarray(e))
    assert _sha(inb) == rec["keygen_in_sha256"]
    assert [int(v) for v in sk[:4]] == rec["sk_head"]
    assert _sha(sk) == rec["sk_sha256"] and _sha(pk) == rec["pk_sha256"] and _sha(temp) == rec["temp_sha256"]
    own = oracle.gaussian_samples(inb[n + 8 * R.r * n: n + 8 * R.r * n + 4 * n].view(np.uint32))
    assert np.count_nonzero(own != e) <= 3
    m = oracle.fill_uniform(n, params.T, 0xC0FFEE)
    c, ee = oracle.encryption_rns(R, pk, m, e0_samples=np.ascontiguousarray(draws[f"{name}_enc_e0"].astype(np.int32)),
                                  e1_samples=np.ascontiguousarray(draws[f"{name}_enc_e1"].astype(np.int32)))
    assert [int(v) for v in c[:4]] == rec["c_head"]
    assert _sha(c) == rec["c_sha256"] and _sha(ee) == rec["e_sha256"]
    plain, _ = oracle.decryption_rns(R, c, sk)
    assert _sha(plain) == rec["plain_sha256"] and np.array_equal(plain, m)


def test_ternary_thresholds_equal_the_float_formula(oracle):
    """csrc/modarith.cuh: ternary_value() counts thresholds (85, 170, 255) instead of evaluating int(float(b) / (255.0f/3)) - 1
    (bfv_keygen.cuh:18-30).  All 256 bytes: the kernel source's two formulas (compiled for the CPU emulator), numpy float32, and the
    oracle's converter agree."""
    from tests import emu
    lib = emu.lib()
    b = np.arange(256, dtype=np.uint8)
    f32 = (b.astype(np.float32) / (np.float32(255.0) / np.float32(3))).astype(np.int32) - 1
    got_float = np.array([lib.emu_ternary(0, int(x)) for x in b])
    got_thr = np.array([lib.emu_ternary(1, int(x)) for x in b])
    assert np.array_equal(got_float, f32) and np.array_equal(got_thr, f32)
    q = 1000
    orc = oracle.ternary_dist_xq(b, 256, [q]).astype(np.int64)      # the _xq formula (the legacy convert_ternary differs: distributions.cuh:204)
    assert np.array_equal(np.where(orc == q - 1, -1, orc), f32)
    assert sorted(set(f32.tolist())) == [-1, 0, 1, 2] and f32[255] == 2 and f32[254] == 1 and f32[85] == 0 and f32[84] == -1
