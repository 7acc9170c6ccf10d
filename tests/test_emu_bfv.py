"""The fused BFV / pointwise / sampling kernel SOURCES executed on the CPU emulator, bit-exact against the oracle.
The gaussian draw is the one value the CPU cannot reproduce bit for bit (normcdfinvf), so the draws the kernels made
are handed to the oracle as an override and everything downstream of them is compared exactly."""
import numpy as np
import pytest

from nttb200 import params
from tests import emu


@pytest.fixture(scope="module")
def ring4k(oracle):
    n, q, roots = params.RNS_SETS["4k_3q"]
    R = oracle.Ring(n, q, roots)
    return R, emu.EmuRing(R)


def test_keystream(oracle):
    ks = emu.sampling(0, None, 40, u0=1, out_size=40 * 8).view(np.uint8)[:40 * 64]
    assert np.array_equal(ks, oracle.generate_random_default(40 * 64))
    # two streams, nonces 5 and 6
    two = emu.sampling(0, None, 8, u0=2, nonce=5, stride=8 * 64, out_size=2 * 8 * 8).view(np.uint8)
    assert np.array_equal(two[:512], oracle.salsa20_keystream(512, b"\x01" * 32, 5))
    assert np.array_equal(two[512:1024], oracle.salsa20_keystream(512, b"\x01" * 32, 6))


def test_converters(oracle):
    q = [274877562881, 274877202433]
    n = 512
    inb = oracle.generate_random_default(64 * 200)
    assert np.array_equal(emu.sampling(1, inb, 2 * n, u0=n, q=q), oracle.ternary_dist_xq(inb, n, q))
    assert np.array_equal(emu.sampling(2, inb, 2 * n, u0=n, q=q), oracle.uniform_dist_xq(inb, n, q))
    allb = np.arange(256, dtype=np.uint8)
    assert np.array_equal(emu.sampling(3, allb, 256, q=q[:1]), oracle.convert_ternary(allb, q[0]))
    w = inb[:8 * 300].view(np.uint64)
    assert np.array_equal(emu.sampling(4, w, 300, q=q[:1]), oracle.convert_range(w, q[0]))


def test_pointwise_kernels(oracle):
    q = 36028797017456641
    qb, mu = oracle.qbit(q), oracle.mu(q)
    n = 1000
    a = oracle.fill_uniform(n, q, 1)
    b = oracle.fill_uniform(n, q, 2)
    a[0], b[0] = q - 1, q - 1
    a[1], b[1] = 0, 5
    a[2], b[2] = q - 3, 3            # a + b == q: the `>` quirk leaves q
    assert np.array_equal(emu.pointwise(0, a, b, s0=q, s1=mu, i0=qb)[1], oracle.barrett(a, b, q))
    assert np.array_equal(emu.pointwise(1, a, s0=q, s1=mu, s2=12345678901, i0=qb)[0], oracle.barrett_int(a, 12345678901, q))
    assert np.array_equal(emu.pointwise(2, a, s0=1024, s2=977)[0], oracle.mod_t(a, 977, 1024))
    got = emu.pointwise(3, a, b, s0=q)[0]
    assert np.array_equal(got, oracle.poly_add(a, b, q)) and int(got[2]) == q
    assert np.array_equal(emu.pointwise(4, a, s0=q, s2=3)[0], oracle.poly_add_integer(a, 3, q))
    assert np.array_equal(emu.pointwise(5, a, b, s0=q)[0], oracle.poly_sub(a, b, q))
    assert np.array_equal(emu.pointwise(6, a, s0=q)[0], oracle.poly_negate(a, q))
    qs = [36028797017456641, 36028797014704129, 18014398506729473]
    qv, muv, qbv = qs, [oracle.mu(x) for x in qs], [oracle.qbit(x) for x in qs]
    A = np.concatenate([oracle.fill_uniform(256, qs[p % 3], 10 + p) for p in range(6)])
    B = np.concatenate([oracle.fill_uniform(256, qs[p % 3], 20 + p) for p in range(6)])
    assert np.array_equal(emu.pointwise(7, A, B, u0=256, u1=3, qv=qv, muv=muv, qbitv=qbv)[1], oracle.barrett_batch(A, B, 256, 6, 3, qv, muv, qbv))
    assert np.array_equal(emu.pointwise(11, A[:768], B[:768], u0=256, qv=qv)[0], oracle.poly_add_negate_xq(A[:768], B[:768], 256, qs))


def test_base_conversion_and_rounding(oracle, ring4k):
    R, _ = ring4k
    n, rp = 256, R.r - 1
    x = np.concatenate([oracle.fill_uniform(n, int(R.q[l]), 40 + l) for l in range(rp)])
    exp = oracle.fast_convert(x, n, rp, R.t, R.gamma, R.gamma_bits, R.mu_gamma, R.bcm)
    got = emu.pointwise(8, x, R.bcm, n=n, s0=R.t, s1=R.gamma, s2=R.mu_gamma, i0=R.gamma_bits, u0=rp, out_size=2 * n)[1]
    assert np.array_equal(got, exp)
    got = emu.pointwise(9, exp, n=n, s0=R.t, s1=R.gamma, s2=R.gamma_div_2, out_size=n)[1]
    assert np.array_equal(got, oracle.dec_round(exp, n, R.t, R.gamma, R.gamma_div_2))
    # divide_and_round_q_last_inplace_loop
    q0, ql = int(R.q[0]), int(R.q[-1])
    a = oracle.fill_uniform(n, q0, 3)
    last = oracle.fill_uniform(n, ql, 4)
    half_mod = (ql >> 1) % q0
    got = emu.pointwise(10, a, last, s0=q0, s1=int(R.mu[0]), s2=half_mod, i0=int(R.qbit[0]), aux=[int(R.inv_q_last_mod_q[0])])[0]
    assert np.array_equal(got, oracle.divide_and_round_q_last_inplace_loop(a, last, q0, half_mod, int(R.inv_q_last_mod_q[0]), int(R.mu[0]), int(R.qbit[0])))


@pytest.mark.parametrize("barrett", [1, 0])
def test_keygen_encrypt_decrypt_vs_oracle(oracle, ring4k, barrett):
    R, er = ring4k
    n, r = R.n, R.r
    sk, pk, es = emu.bfv(0, er, barrett)
    osk, opk, otemp, oin = oracle.keygen_rns(R, e_samples=np.ascontiguousarray(es[0]))
    assert np.array_equal(sk, osk) and np.array_equal(pk, opk)
    # the CPU inverse-normal and the emulator's agree except (rarely) at a truncation boundary
    assert np.count_nonzero(oracle.gaussian_samples(oin[n + 8 * r * n: n + 8 * r * n + 4 * n].view(np.uint32)) != es[0]) <= 2
    m = oracle.fill_uniform(n, R.t, 0xBEEF)
    c, es2 = emu.bfv(1, er, barrett, pk=pk, m=m)
    oc, oe = oracle.encryption_rns(R, pk, m, e0_samples=np.ascontiguousarray(es2[0, 0]), e1_samples=np.ascontiguousarray(es2[0, 1]))
    assert np.array_equal(c, oc)
    out, _ = emu.bfv(2, er, barrett, sk=sk, c=c)
    oplain, _ = oracle.decryption_rns(R, c, sk)
    assert np.array_equal(out[0], oplain) and np.array_equal(out[0], m)


def test_decrypt_kat_on_emulator(oracle, ring4k):
    """The reference's golden vector through the fused decryption kernels (both NTT flavours)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "decryption_kat.npz"))
    R, er = ring4k
    sk = np.zeros(R.r * R.n, dtype=np.uint64)
    sk[:8192] = g["sk_host"]
    for barrett in (1, 0):
        out, _ = emu.bfv(2, er, barrett, sk=sk, c=g["c_host"])
        assert np.array_equal(out[0], np.arange(4096, dtype=np.uint64) % 10)


def test_batched_items_use_their_own_nonce(oracle, ring4k):
    R, er = ring4k
    n, r = R.n, R.r
    B = 2
    sk, pk, es = emu.bfv(0, er, 0, batch=B, nonce0=7)
    for k in range(B):
        oracle.set_nonce(7 + k)
        osk, opk, _, _ = oracle.keygen_rns(R, e_samples=np.ascontiguousarray(es[k]))
        assert np.array_equal(sk[k * r * n:(k + 1) * r * n], osk) and np.array_equal(pk[k * 2 * r * n:(k + 1) * 2 * r * n], opk)
    oracle.set_nonce(0)
    m = np.concatenate([oracle.fill_uniform(n, R.t, 100 + k) for k in range(B)])
    c, _ = emu.bfv(1, er, 0, batch=B, nonce0=3, pk=pk, m=m, per_item_keys=1)
    out, _ = emu.bfv(2, er, 0, batch=B, sk=sk, c=c, per_item_keys=1)
    assert np.array_equal(out.reshape(-1), m)


@pytest.mark.parametrize("ops", [(3, 4), (5, 6)], ids=["lazy", "general"])
def test_fused_key_product_matches_oracle(oracle, ring4k, ops):
    """Loaded-key path: strided pass, ONE fused kernel (contig NTT pass (.) key by Shoup companion, contig INTT pass), strided pass.
    Same ciphertext / plaintext bits as the oracle (and so as the unfused kernels), batch of 2 sharing the key, plus the KAT."""
    import os
    R, er = ring4k
    n, r = R.n, R.r
    enc, dec = ops
    sk, pk, _ = emu.bfv(0, er, 0)
    B = 2
    m = np.concatenate([oracle.fill_uniform(n, R.t, 0xF00 + k) for k in range(B)])
    c, es = emu.bfv(enc, er, 0, batch=B, nonce0=11, pk=pk, m=m)
    for k in range(B):
        oracle.set_nonce(11 + k)
        oc, _ = oracle.encryption_rns(R, pk, m[k * n:(k + 1) * n], e0_samples=np.ascontiguousarray(es[k, 0]),
                                      e1_samples=np.ascontiguousarray(es[k, 1]))
        assert np.array_equal(c[k * 2 * r * n:(k + 1) * 2 * r * n], oc)
    oracle.set_nonce(0)
    out, _ = emu.bfv(dec, er, 0, batch=B, sk=sk, c=c)
    assert np.array_equal(out.reshape(-1), m)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "decryption_kat.npz"))
    skk = np.zeros(r * n, dtype=np.uint64)
    skk[:8192] = g["sk_host"]
    out, _ = emu.bfv(dec, er, 0, sk=skk, c=g["c_host"])
    assert np.array_equal(out[0], np.arange(4096, dtype=np.uint64) % 10)


def test_all_lazy_epilogues_match_oracle(oracle):
    """The ALL_LAZY encryption epilogue / ALL_FAST decryption epilogue (the instantiations nttb200_bfv_create selects when every
    limb's Barrett is provably exact and q_last <= 2 q_i) against the oracle on a parameter set that qualifies (8k_4q), both
    the unfused and the fused-key pipelines; message coefficients at the extremes 0 and t-1 included."""
    n, q, roots = params.RNS_SETS["8k_4q"]
    R = oracle.Ring(n, q, roots)
    er = emu.EmuRing(R)
    sk, pk, _ = emu.bfv(0, er, 0)
    m = oracle.fill_uniform(n, R.t, 0xFA57)
    m[:4] = [0, R.t - 1, 1, R.t - 1]
    emu.lib().emu_set_epilogue_fast(1)
    try:
        for enc, dec in ((1, 2), (3, 4)):
            c, es2 = emu.bfv(enc, er, 0, pk=pk, m=m)
            oc, _ = oracle.encryption_rns(R, pk, m, e0_samples=np.ascontiguousarray(es2[0, 0]), e1_samples=np.ascontiguousarray(es2[0, 1]))
            assert np.array_equal(c, oc)
            out, _ = emu.bfv(dec, er, 0, sk=sk, c=c)
            assert np.array_equal(out[0], m)
    finally:
        emu.lib().emu_set_epilogue_fast(0)


def test_wire_format_and_homomorphic_helper_kernels_on_emulator(oracle):
    """k_ct_pack / k_ct_unpack (format definition in include/nttb200.h, restated here with numpy), k_ct_add and k_plain_lift."""
    import ctypes as C
    n, qs, _ = params.RNS_SETS["4k_3q"]
    r, batch, t = len(qs), 2, params.T
    qa = np.array(qs, dtype=np.uint64)
    qbit = np.array([int(q).bit_length() for q in qs], dtype=np.uint32)
    woff = np.zeros(r - 1, dtype=np.uint32)
    for l in range(1, r - 1):
        woff[l] = woff[l - 1] + n // 64 * qbit[l - 1]
    half_words = int(woff[-1] + n // 64 * qbit[r - 2])
    c = np.concatenate([oracle.fill_uniform(n, qs[l], 0x77 + 10 * k + l) for k in range(batch * 2) for l in range(r)])
    c[0], c[1] = 0, qs[0] - 1
    packed = np.zeros(batch * 2 * half_words, dtype=np.uint64)
    u, u32 = C.c_ulonglong, C.c_uint
    lib = emu.lib()
    assert lib.emu_ct_ops(0, emu.p(c, u), emu.p(packed, u), n, r, batch, emu.p(qa, u), emu.p(qbit, u32), emu.p(woff, u32), half_words, C.c_ulonglong(t)) == 0
    hc = c.reshape(batch * 2, r, n)
    exp = []
    for kh in range(batch * 2):
        for l in range(r - 1):
            bits = ((hc[kh, l][:, None] >> np.arange(int(qbit[l]), dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1)
            exp.append(np.packbits(bits, bitorder="little").view(np.uint64))
    assert np.array_equal(packed, np.concatenate(exp))
    back = np.full_like(c, 0xFFFFFFFFFFFFFFFF)
    assert lib.emu_ct_ops(1, emu.p(back, u), emu.p(packed, u), n, r, batch, emu.p(qa, u), emu.p(qbit, u32), emu.p(woff, u32), half_words, C.c_ulonglong(t)) == 0
    hb = back.reshape(batch * 2, r, n)
    assert np.array_equal(hb[:, :r - 1], hc[:, :r - 1]) and (hb[:, r - 1] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()   # padding limb is the caller's
    # limb-wise addition on the stored limbs
    d = np.concatenate([oracle.fill_uniform(n, qs[l], 0x99 + 10 * k + l) for k in range(batch * 2) for l in range(r)])
    s = c.copy()
    assert lib.emu_ct_ops(2, emu.p(s, u), emu.p(d, u), n, r, batch, emu.p(qa, u), emu.p(qbit, u32), emu.p(woff, u32), half_words, C.c_ulonglong(t)) == 0
    hs, hd = s.reshape(batch * 2, r, n), d.reshape(batch * 2, r, n)
    for l in range(r - 1):
        assert np.array_equal(hs[:, l], (hc[:, l] + hd[:, l]) % np.uint64(qs[l]))
    assert np.array_equal(hs[:, r - 1], hc[:, r - 1])
    # centred plaintext lift
    m = oracle.fill_uniform(batch * n, 1 << 20, 0x4242)                      # deliberately not reduced mod t
    P = np.zeros(batch * (r - 1) * n, dtype=np.uint64)
    assert lib.emu_ct_ops(3, emu.p(P, u), emu.p(m, u), n, r, batch, emu.p(qa, u), emu.p(qbit, u32), emu.p(woff, u32), half_words, C.c_ulonglong(t)) == 0
    mm = (m % np.uint64(t)).reshape(batch, n).astype(np.int64)
    cen = np.where(mm <= t // 2, mm, mm - t)
    for l in range(r - 1):
        assert np.array_equal(P.reshape(batch, r - 1, n)[:, l], (cen % qs[l]).astype(np.uint64))
