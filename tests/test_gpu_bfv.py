"""GPU parity of the pointwise kernels, the sampler and the BFV pipelines through the C ABI against the CPU oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from nttb200 import params  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "decryption_kat.npz")


def _dev_ring(oracle, name):
    from tests.gpu_util import to_dev
    n, q, roots = params.RNS_SETS[name]
    R = oracle.Ring(n, q, roots)
    d = dict(q=to_dev(R.qa), mu=to_dev(R.mu), qbit=to_dev(R.qbit), psi=to_dev(R.psi.reshape(-1)), psiinv=to_dev(R.psiinv.reshape(-1)),
             iql=to_dev(R.inv_q_last_mod_q), qdt=to_dev(R.qi_div_t), ptg=to_dev(R.prod_t_gamma_mod_q), ipq=to_dev(R.inv_punctured_q),
             bcm=to_dev(R.bcm))
    return R, d


def test_keystream_and_converters(oracle):
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    nbytes = 64 * 1000 + 17        # trailing partial block is dropped (NBLKS = n / 64)
    buf = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    nttb200.generate_random_default(buf, nbytes)
    ks = to_host(buf)
    assert np.array_equal(ks[:64000], oracle.generate_random_default(64000)) and not ks[64000:].any()
    # generate_random: key 0x4D in bytes 0..23, tail left by the previous generate_random_default call = 0x01
    nttb200.generate_random(buf, 6400)
    assert np.array_equal(to_host(buf)[:6400], oracle.generate_random(6400, b"\x01" * 8))
    # explicit key / nonce / two streams
    two = torch.zeros(2 * 640, dtype=torch.uint8, device="cuda")
    nttb200.salsa20_keystream(two, 10, 2, 640, bytes(range(32)), 41)
    got = to_host(two)
    assert np.array_equal(got[:640], oracle.salsa20_keystream(640, bytes(range(32)), 41))
    assert np.array_equal(got[640:], oracle.salsa20_keystream(640, bytes(range(32)), 42))
    # converters
    q = [274877562881, 274877202433, 274877153281]
    qd = to_dev(np.array(q, dtype=np.uint64))
    n = 2048
    inb = oracle.generate_random_default(9 * 3 * n + 4 * n)
    ind = to_dev(inb)
    out = torch.zeros(3 * n, dtype=torch.int64, device="cuda")
    nttb200.ternary_dist_xq(ind, out, n, 3, qd)
    assert np.array_equal(to_host(out), oracle.ternary_dist_xq(inb, n, q))
    nttb200.uniform_dist_xq(ind[n:], out, n, 3, qd)
    assert np.array_equal(to_host(out), oracle.uniform_dist_xq(inb[n:], n, q))
    nttb200.gaussian_dist_xq(ind[n + 8 * 3 * n:], out, n, 3, qd)
    got = to_host(out)
    exp = oracle.gaussian_dist_xq(inb[n + 8 * 3 * n:], n, q)
    # normcdfinvf cannot be reproduced bit for bit on the CPU: allow a handful of +-1 draws at truncation boundaries
    bad = np.nonzero(got != exp)[0]
    assert bad.size <= 3
    allb = to_dev(np.arange(256, dtype=np.uint8))
    o256 = torch.zeros(256, dtype=torch.int64, device="cuda")
    nttb200.ternary_dist(allb, o256, 256, None, q[0])
    assert np.array_equal(to_host(o256), oracle.convert_ternary(np.arange(256, dtype=np.uint8), q[0]))
    w = inb[:8 * 512].view(np.uint64)
    o512 = torch.zeros(512, dtype=torch.int64, device="cuda")
    nttb200.uniform_dist(to_dev(w), o512, 512, None, q[1])
    assert np.array_equal(to_host(o512), oracle.convert_range(w, q[1]))


def test_pointwise_kernels(oracle):
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    q = 36028797017456641
    qb, mu = oracle.qbit(q), oracle.mu(q)
    n = 4096 + 300     # not a multiple of 256: the reference's <<<n/256,256>>> would drop the tail, grid-stride does not
    a = oracle.fill_uniform(n, q, 1)
    b = oracle.fill_uniform(n, q, 2)
    a[0], b[0] = q - 1, q - 1
    a[2], b[2] = q - 3, 3
    bd = to_dev(b)

    def run(fn, *args):
        d = to_dev(a)
        fn(d, *args)
        return to_host(d)

    assert np.array_equal(run(nttb200.barrett, bd, n, q, mu, qb), oracle.barrett(a, b, q))
    assert np.array_equal(run(nttb200.poly_mul_int, 987654321, n, None, q, mu, qb), oracle.barrett_int(a, 987654321, q))
    assert np.array_equal(run(nttb200.poly_mul_int_t, 977, n, None, 1024), oracle.mod_t(a, 977, 1024))
    got = run(nttb200.poly_add_device, bd, n, None, q)
    assert np.array_equal(got, oracle.poly_add(a, b, q)) and int(got[2]) == q
    assert np.array_equal(run(nttb200.poly_add_integer_device, 3, n, None, q), oracle.poly_add_integer(a, 3, q))
    assert np.array_equal(run(nttb200.poly_sub_device, bd, n, None, q), oracle.poly_sub(a, b, q))
    assert np.array_equal(run(nttb200.poly_negate_device, n, None, q), oracle.poly_negate(a, q))
    R, d = _dev_ring(oracle, "4k_3q")
    nn, rp = 4096, 2
    x = np.concatenate([oracle.fill_uniform(nn, int(R.q[l]), 40 + l) for l in range(rp)])
    res = torch.zeros(2 * nn, dtype=torch.int64, device="cuda")
    nttb200.fast_convert_array_kernels(to_dev(x), res, R.t, d["bcm"], rp, R.gamma, R.gamma_bits, R.mu_gamma, nn)
    exp = oracle.fast_convert(x, nn, rp, R.t, R.gamma, R.gamma_bits, R.mu_gamma, R.bcm)
    assert np.array_equal(to_host(res), exp)
    o = torch.zeros(nn, dtype=torch.int64, device="cuda")
    nttb200.dec_round(res, o, R.t, R.gamma, R.gamma_div_2, nn)
    assert np.array_equal(to_host(o), oracle.dec_round(exp, nn, R.t, R.gamma, R.gamma_div_2))
    A = np.concatenate([oracle.fill_uniform(nn, int(R.q[p % 3]), 10 + p) for p in range(6)])
    B = np.concatenate([oracle.fill_uniform(nn, int(R.q[p % 3]), 20 + p) for p in range(6)])
    Ad = to_dev(A)
    nttb200.barrett_batch(Ad, to_dev(B), nn, 6, 3, d["q"], d["mu"], d["qbit"])
    assert np.array_equal(to_host(Ad), oracle.barrett_batch(A, B, nn, 6, 3, R.qa, R.mu, R.qbit))


def test_decryption_kat_both_front_ends(oracle):
    """decryption_test.cu's golden vector: c_host / sk_host -> i % 10, through the stateless reference-contract call
    (plaintext inside c, demo.cu:299) and through the batched context API."""
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    g = np.load(GOLD)
    R, d = _dev_ring(oracle, "4k_3q")
    n, rp = 4096, 2
    expect = np.arange(n, dtype=np.uint64) % 10
    c = to_dev(g["c_host"])
    sk = to_dev(g["sk_host"])
    nttb200.decryption_rns(c, sk, d["psi"], d["psiinv"], n, rp, d["bcm"], R.t, R.gamma, R.mu_gamma, R.gamma_bits, int(R.neg_inv[0]),
                           int(R.neg_inv[1]), R.gamma_div_2, d["q"], d["mu"], d["qbit"], d["ipq"], d["ptg"])
    assert np.array_equal(to_host(c)[n * (rp - 1): n * rp], expect)
    bfv = nttb200.Bfv(n, R.q, R.psi_roots)
    B = 3
    cb = to_dev(np.tile(g["c_host"], B))
    out = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    bfv.decrypt(out, cb, sk, batch=B)
    assert np.array_equal(to_host(out), np.tile(expect, B))
    bfv.close()


@pytest.mark.parametrize("name", ["4k_3q", "8k_4q", "16k_5q", "32k_9q"])
def test_pipelines_vs_oracle(oracle, name):
    """keygen -> encrypt -> decrypt, stateless front end, every buffer compared with the oracle (gaussian draws taken
    from the GPU: the first n / 2n words of temp / e, see include/nttb200.h)."""
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    R, d = _dev_ring(oracle, name)
    n, r = R.n, R.r
    rn = r * n
    inb = torch.zeros(9 * rn + 4 * n, dtype=torch.uint8, device="cuda")
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    temp = torch.zeros(rn, dtype=torch.int64, device="cuda")
    nttb200.keygen_rns(inb, r, n, sk, pk, temp, d["psi"], d["psiinv"], d["q"], d["mu"], d["qbit"])
    es = to_host(temp).view(np.int32)[:n].copy()
    osk, opk, _, oin = oracle.keygen_rns(R, e_samples=es)
    assert np.array_equal(to_host(inb), oin)
    assert np.array_equal(to_host(sk), osk) and np.array_equal(to_host(pk), opk)
    cpu_draws = oracle.gaussian_samples(oin[n + 8 * rn: n + 8 * rn + 4 * n].view(np.uint32))
    assert np.count_nonzero(cpu_draws != es) <= 3 and np.abs(cpu_draws - es).max() <= 1

    m = oracle.fill_uniform(n, R.t, 0xC0FFEE)
    c = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    e = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    nttb200.encryption_rns(c, pk, inb, e, n, d["psi"], d["psiinv"], to_dev(m), d["qdt"], R.t, r, d["q"], d["mu"], d["qbit"], d["iql"])
    es2 = to_host(e).view(np.int32)[:2 * n].copy()
    oc, _ = oracle.encryption_rns(R, opk, m, e0_samples=np.ascontiguousarray(es2[:n]), e1_samples=np.ascontiguousarray(es2[n:]))
    assert np.array_equal(to_host(c), oc)

    nttb200.decryption_rns(c, sk, d["psi"], d["psiinv"], n, r - 1, d["bcm"], R.t, R.gamma, R.mu_gamma, R.gamma_bits, int(R.neg_inv[0]),
                           int(R.neg_inv[1]), R.gamma_div_2, d["q"], d["mu"], d["qbit"], d["ipq"], d["ptg"])
    plain = to_host(c)[n * (r - 2): n * (r - 1)]
    assert np.array_equal(plain, m)
    oplain, _ = oracle.decryption_rns(R, oc, osk)
    assert np.array_equal(plain, oplain)


def test_batched_context_pipelines(oracle):
    """Batched API (Shoup NTT): item k == the oracle run with nonce nonce0 + k; full batch round-trips."""
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    name = "8k_3q"
    R, d = _dev_ring(oracle, name)
    n, r = R.n, R.r
    rn = r * n
    B = 5
    bfv = nttb200.Bfv(n, R.q, R.psi_roots)
    sk = torch.zeros(B * rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk, batch=B, nonce0=11)
    hsk, hpk = to_host(sk), to_host(pk)
    # reference draws for item k come from the stateless sampler on the same keystream
    for k in (0, 4):
        oracle.set_nonce(11 + k)
        ks = oracle.generate_random_default(9 * rn + 4 * n)
        t = torch.zeros(n, dtype=torch.int64, device="cuda")
        nttb200.gaussian_dist_xq(to_dev(ks[n + 8 * rn:]), t, n, 1, d["q"])
        v = to_host(t).astype(np.int64)
        es = np.where(v > int(R.q[0]) // 2, v - int(R.q[0]), v).astype(np.int32)
        osk, opk, _, _ = oracle.keygen_rns(R, e_samples=es)
        assert np.array_equal(hsk[k * rn:(k + 1) * rn], osk) and np.array_equal(hpk[k * 2 * rn:(k + 1) * 2 * rn], opk)
    oracle.set_nonce(0)
    m = np.concatenate([oracle.fill_uniform(n, R.t, 300 + k) for k in range(B)])
    c = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    bfv.encrypt(c, pk, to_dev(m), batch=B, nonce0=100, pk_per_item=True)
    out = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    bfv.decrypt(out, c, sk, batch=B, sk_per_item=True)
    assert np.array_equal(to_host(out), m)
    # one shared key pair for a batch of ciphertexts
    c2 = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    bfv.encrypt(c2, pk[:2 * rn], to_dev(m), batch=B, nonce0=0)
    bfv.decrypt(out, c2, sk[:rn], batch=B)
    assert np.array_equal(to_host(out), m)
    bfv.close()


def test_pipelines_are_cuda_graph_capturable(oracle):
    """Single-item calls are launch-bound; after reserve() nothing on the path allocates or synchronises, so the whole
    encrypt / decrypt call captures into a CUDA graph and replays to the same bits."""
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = params.RNS_SETS["16k_5q"]
    r = len(qs)
    rn = r * n
    bfv = nttb200.Bfv(n, qs, roots)
    bfv.reserve(1)
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk)
    m = to_dev(oracle.fill_uniform(n, params.T, 31337))
    c_ref = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.encrypt(c_ref, pk, m)                      # eager (also warms up attributes / tensor-map entry point)
    out_ref = torch.zeros(n, dtype=torch.int64, device="cuda")
    bfv.decrypt(out_ref, c_ref.clone(), sk)
    c = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    out = torch.zeros(n, dtype=torch.int64, device="cuda")
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            bfv.encrypt(c, pk, m, stream=s)
            bfv.decrypt(out, c, sk, stream=s)
    for _ in range(3):
        c.zero_()
        out.zero_()
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, m) and torch.equal(out, out_ref)
    bfv.close()


@pytest.mark.parametrize("bits,logn,r", [(57, 11, 3), (38, 12, 3), (57, 13, 4), (57, 14, 3), (50, 15, 5), (57, 16, 2), (57, 17, 2),
                                         (60, 11, 3), (60, 13, 2), (59, 16, 2)])
def test_loaded_key_fused_path(oracle, bits, logn, r):
    """nttb200_bfv_load_keys + NULL key pointers route the key products through the fused "contig NTT pass (.) key -> contig INTT pass"
    kernel.  Ciphertexts and plaintexts must equal the unfused pipelines' bit for bit (those are pinned to the oracle above), for
    every schedule 2^11..2^17 and for both arithmetic families (q < 2^57 lazy, q < 2^62 general)."""
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n = 1 << logn
    qs, roots = params.find_ntt_primes(bits, n, r)
    rn = r * n
    B = 3
    bfv = nttb200.Bfv(n, qs, roots)
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk, nonce0=5)
    bfv.load_keys(sk, pk)
    m = to_dev(np.concatenate([oracle.fill_uniform(n, params.T, 900 + k) for k in range(B)]))
    c_plain = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    c_fused = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    bfv.encrypt(c_plain, pk, m, batch=B, nonce0=77)
    bfv.encrypt(c_fused, None, m, batch=B, nonce0=77)
    assert torch.equal(c_plain, c_fused)
    out_plain = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    out_fused = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    bfv.decrypt(out_plain, c_plain, sk, batch=B)
    bfv.decrypt(out_fused, c_fused, None, batch=B)
    assert torch.equal(out_fused, m) and torch.equal(out_plain, m)
    assert torch.equal(c_plain, c_fused)          # decryption leaves the same c1 scratch behind in both paths
    if logn <= 13:                                # and against the oracle directly where it is quick
        R = oracle.Ring(n, qs, roots)
        fresh = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
        bfv.encrypt(fresh, None, m, batch=1, nonce0=77)
        oplain, _ = oracle.decryption_rns(R, to_host(fresh), to_host(sk))
        assert np.array_equal(oplain, to_host(m)[:n])
    bfv.close()


def test_loaded_key_kat(oracle):
    """The reference's decryption golden vector through the fused path."""
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    g = np.load(GOLD)
    n, qs, roots = params.RNS_SETS["4k_3q"]
    bfv = nttb200.Bfv(n, qs, roots)
    sk = np.zeros(3 * n, dtype=np.uint64)
    sk[:8192] = g["sk_host"]
    bfv.load_keys(to_dev(sk), None)
    out = torch.zeros(n, dtype=torch.int64, device="cuda")
    bfv.decrypt(out, to_dev(g["c_host"]), None)
    assert np.array_equal(to_host(out), np.arange(4096, dtype=np.uint64) % 10)
    with pytest.raises(Exception):
        bfv.encrypt(to_dev(g["c_host"]), None, out)     # no public key loaded: refused, not silently computed
    bfv.close()


@pytest.mark.parametrize("name,batch", [("4k_3q", 3), ("8k_4q", 2), ("32k_16q", 2)])
def test_homomorphic_add_and_plain_multiply(oracle, name, batch):
    """SURVEY.md 8f-4 (the reference stops at decryption): Dec(Enc(m1) + Enc(m2)) == m1 + m2 mod t and
    Dec(Enc(m) * p) == m * p mod (X^n + 1, t), the latter against the schoolbook negacyclic product of helper.h:95-126 (oracle) and,
    independently, against the oracle's own decryption of the product ciphertext; shared and per-item plaintext factors."""
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = params.RNS_SETS[name]
    R = oracle.Ring(n, qs, roots)
    t, rn = params.T, len(qs) * n
    bfv = nttb200.Bfv(n, qs, roots)
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk)
    m1 = np.concatenate([oracle.fill_uniform(n, t, 0x111 + k) for k in range(batch)])
    m2 = np.concatenate([oracle.fill_uniform(n, t, 0x222 + k) for k in range(batch)])
    m1[:3] = [0, t - 1, t - 1]
    m2[:3] = [t - 1, t - 1, 1]
    c1 = torch.zeros(batch * 2 * rn, dtype=torch.int64, device="cuda")
    c2 = torch.zeros_like(c1)
    bfv.encrypt(c1, pk, to_dev(m1), batch=batch, nonce0=5)
    bfv.encrypt(c2, pk, to_dev(m2), batch=batch, nonce0=50)
    out = torch.zeros(batch * n, dtype=torch.int64, device="cuda")
    # addition
    s = c1.clone()
    bfv.add(s, c2, batch=batch)
    bfv.decrypt(out, s.clone(), sk, batch=batch)
    assert np.array_equal(to_host(out), (m1 + m2) % np.uint64(t))
    # plaintext addition (shared plaintext, then one per item)
    s = c1.clone()
    bfv.add_plain(s, to_dev(m2[:n]), batch=batch)
    bfv.decrypt(out, s, sk, batch=batch)
    assert np.array_equal(to_host(out), (m1 + np.tile(m2[:n], batch)) % np.uint64(t))
    s = c1.clone()
    bfv.add_plain(s, to_dev(m2), batch=batch, plain_per_item=True)
    bfv.decrypt(out, s, sk, batch=batch)
    assert np.array_equal(to_host(out), (m1 + m2) % np.uint64(t))
    # plaintext multiplication: a sparse factor keeps the schoolbook reference cheap; one shared factor, then one per item
    def negacyclic_mod_t(a, b):
        res = np.zeros(n, dtype=np.int64)
        for i in np.nonzero(b)[0]:
            v = int(b[i])
            res[i:] += v * a[:n - i].astype(np.int64)
            res[:i] -= v * a[n - i:].astype(np.int64)
        return (res % t).astype(np.uint64)
    p_shared = np.zeros(n, dtype=np.uint64)
    p_shared[[0, 1, n // 2, n - 1]] = [3, t - 1, 7, 512]
    prod = c1.clone()
    bfv.mul_plain(prod, to_dev(p_shared), batch=batch)
    host_prod = to_host(prod).copy()
    bfv.decrypt(out, prod, sk, batch=batch)
    got = to_host(out)
    for k in range(batch):
        assert np.array_equal(got[k * n:(k + 1) * n], negacyclic_mod_t(m1[k * n:(k + 1) * n], p_shared)), f"item {k}"
    # the oracle's decryption (CPU restatement of decryption_rns) agrees on the product ciphertext
    plain, _ = oracle.decryption_rns(R, host_prod[:2 * rn], to_host(sk))
    assert np.array_equal(plain, got[:n])
    p_items = np.zeros(batch * n, dtype=np.uint64)
    for k in range(batch):
        p_items[k * n + k] = 2 + k
        p_items[k * n + n - 1 - k] = t - 3
    prod = c2.clone()
    bfv.mul_plain(prod, to_dev(p_items), batch=batch, plain_per_item=True)
    bfv.decrypt(out, prod, sk, batch=batch)
    got = to_host(out)
    for k in range(batch):
        assert np.array_equal(got[k * n:(k + 1) * n], negacyclic_mod_t(m2[k * n:(k + 1) * n], p_items[k * n:(k + 1) * n])), f"item {k}"
    bfv.close()


@pytest.mark.parametrize("name,batch", [("4k_3q", 2), ("16k_5q", 2), ("32k_16q", 3)])
def test_wire_format_round_trip(oracle, name, batch):
    """pack -> bytes match a numpy restatement of the format definition in include/nttb200.h -> unpack restores every
    stored limb bit for bit (padding limb zeroed) and the unpacked ciphertext still decrypts to the message."""
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = params.RNS_SETS[name]
    r, rn, t = len(qs), len(qs) * n, params.T
    bfv = nttb200.Bfv(n, qs, roots)
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk)
    m = np.concatenate([oracle.fill_uniform(n, t, 0x333 + k) for k in range(batch)])
    c = torch.zeros(batch * 2 * rn, dtype=torch.int64, device="cuda")
    bfv.encrypt(c, pk, to_dev(m), batch=batch, nonce0=9)
    words = bfv.packed_words()
    qbits = [int(q).bit_length() for q in qs[:-1]]
    assert words == 2 * (n // 64) * sum(qbits)
    packed = torch.zeros(batch * words, dtype=torch.int64, device="cuda")
    bfv.pack(packed, c, batch=batch)
    hc, hp = to_host(c).reshape(batch, 2, r, n), to_host(packed).reshape(batch, words)
    for k in range(batch):
        exp = []
        for h in range(2):
            for l, qb in enumerate(qbits):
                bits = ((hc[k, h, l][:, None] >> np.arange(qb, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1)
                exp.append(np.packbits(bits, bitorder="little").view(np.uint64))
        assert np.array_equal(hp[k], np.concatenate(exp)), f"item {k}"
    back = torch.full_like(c, -1)
    bfv.unpack(back, packed, batch=batch)
    hb = to_host(back).reshape(batch, 2, r, n)
    assert np.array_equal(hb[:, :, :r - 1], hc[:, :, :r - 1]) and not hb[:, :, r - 1].any()
    out = torch.zeros(batch * n, dtype=torch.int64, device="cuda")
    bfv.decrypt(out, back, sk, batch=batch)
    assert np.array_equal(to_host(out), m)
    bfv.close()
