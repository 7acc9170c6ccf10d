"""GPU parity of the NTT / INTT through the C ABI (libnttb200.so) against the CPU oracle.  Bit-exact: integer work."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from nttb200 import params  # noqa: E402


def _ring(oracle, logn, limbs):
    n = 1 << logn
    if n in params.GET_PARAMS and limbs == 1:
        q, psi = params.GET_PARAMS[n][:2]
        return n, [q], [psi]
    for name in ("32k_16q", "16k_9q", "8k_4q", "4k_3q"):
        nn, qs, roots = params.RNS_SETS[name]
        if nn == n and limbs <= len(qs):
            return n, qs[:limbs], roots[:limbs]
    qs, roots = params.find_ntt_primes(60, n, limbs)
    return n, qs, roots


def _tables(oracle, n, qs, roots):
    tabs = [oracle.fill_psi_tables(r, q, n) for q, r in zip(qs, roots)]
    return np.stack([t[0] for t in tabs]), np.stack([t[1] for t in tabs])


def _expect(oracle, a, n, qs, psi, psiinv, num, division, inverse, literal=False):
    """literal=False: the exact transform (canonical residues).  literal=True: the reference's arithmetic operation for
    operation (oracle.forward_ntt / inverse_ntt) -- identical except on the rare inputs where the reference's single-
    correction Barrett returns a value in [q, 2q) (only possible for primes with frac(2^(2*qbit)/q) > 3/4)."""
    out = a.copy().reshape(num, n)
    for p in range(num):
        l = p % division
        if literal:
            out[p] = oracle.inverse_ntt(out[p], qs[l], psiinv[l]) if inverse else oracle.forward_ntt(out[p], qs[l], psi[l])
        else:
            out[p] = oracle.inverse_ntt_fast(out[p], qs[l], psiinv[l]) if inverse else oracle.forward_ntt_fast(out[p], qs[l], psi[l])
    return out.reshape(-1)


@pytest.mark.parametrize("tma", [1, 0])
# n <= 4096 has three schedules by polynomial count: <= 32 one cluster per transform (ntt_cluster.cuh), <= 64 one CTA per transform
# (ntt_single_pass), more: the two batched passes -- (11, 2, 40), (12, 3, 33), (12, 2, 70), (11, 1, 32) pin the boundaries.
@pytest.mark.parametrize("logn,limbs,num", [(11, 1, 3), (12, 1, 2), (12, 3, 5), (13, 1, 2), (14, 9, 11), (15, 16, 40), (16, 3, 4), (17, 2, 3),
                                            (11, 2, 40), (12, 3, 33), (12, 2, 70), (11, 1, 32)])
def test_ctx_forward_inverse_vs_oracle(oracle, logn, limbs, num, tma):
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = _ring(oracle, logn, limbs)
    psi, psiinv = _tables(oracle, n, qs, roots)
    ctx = nttb200.Context(n, qs, roots)
    ctx.set_tma(bool(tma))
    a = np.concatenate([oracle.fill_uniform(n, qs[p % limbs], 0x5EED0000 + p) for p in range(num)])
    a[0], a[1], a[2] = 0, 1, qs[0] - 1
    d = to_dev(a)
    ctx.forward_ntt_batch(d, num, limbs)
    fwd = to_host(d)
    assert np.array_equal(fwd, _expect(oracle, a, n, qs, psi, psiinv, num, limbs, False))
    ctx.inverse_ntt_batch(d, num, limbs)
    assert np.array_equal(to_host(d), a)
    ctx.close()


def test_ctx_tables_match_reference_layout(oracle):
    """Context tables generated on the library side == parameter.h:5-12 fillTablePsi128 (oracle restatement)."""
    import nttb200
    n, qs, roots = params.RNS_SETS["8k_3q"]
    psi, psiinv = _tables(oracle, n, qs, roots)
    ctx = nttb200.Context(n, qs, roots)
    assert np.array_equal(nttb200.download(ctx.psi_table, 3 * n).reshape(3, n), psi)
    assert np.array_equal(nttb200.download(ctx.psiinv_table, 3 * n).reshape(3, n), psiinv)
    ctx.close()


@pytest.mark.parametrize("logn,limbs,num", [(11, 1, 1), (12, 3, 6), (13, 3, 3), (15, 9, 18), (16, 2, 2)])
def test_stateless_reference_contract_path(oracle, logn, limbs, num):
    """forwardNTT_batch / inverseNTT_batch given nothing but the reference's tables and q/mu/qbit device arrays."""
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = _ring(oracle, logn, limbs)
    psi, psiinv = _tables(oracle, n, qs, roots)
    qd = to_dev(np.array(qs, dtype=np.uint64))
    mud = to_dev(np.array([oracle.mu(q) for q in qs], dtype=np.uint64))
    qbd = to_dev(np.array([oracle.qbit(q) for q in qs], dtype=np.uint32))
    psid, psiinvd = to_dev(psi.reshape(-1)), to_dev(psiinv.reshape(-1))
    a = np.concatenate([oracle.fill_uniform(n, qs[p % limbs], 77 + p) for p in range(num)])
    d = to_dev(a)
    nttb200.forwardNTT_batch(d, n, psid, num, limbs, qd, mud, qbd)
    fwd = to_host(d)
    assert np.array_equal(fwd, _expect(oracle, a, n, qs, psi, psiinv, num, limbs, False, literal=True))
    nttb200.inverseNTT_batch(d, n, psiinvd, num, limbs, qd, mud, qbd)
    # [12-3-6] contains an input on which the reference's own Barrett glitches (q = 68719230977): this path must
    # reproduce the reference bit for bit even there, so the expectation is the literal restatement, not `a`.
    assert np.array_equal(to_host(d), _expect(oracle, fwd, n, qs, psi, psiinv, num, limbs, True, literal=True))
    # single-polynomial API with explicit constants (forwardNTT / inverseNTT)
    d1 = to_dev(a[:n])
    nttb200.forwardNTT(d1, n, None, qs[0], oracle.mu(qs[0]), oracle.qbit(qs[0]), psid)
    assert np.array_equal(to_host(d1), oracle.forward_ntt(a[:n], qs[0], psi[0]))
    nttb200.inverseNTT(d1, n, None, qs[0], oracle.mu(qs[0]), oracle.qbit(qs[0]), psiinvd)
    assert np.array_equal(to_host(d1), a[:n])


def test_full_size_c2_properties(oracle):
    """BASELINE config 2: 1024 polynomials, N = 2^15, 16 limbs.  Size-independent checks: INTT(NTT(a)) == a,
    linearity NTT(a + b) == NTT(a) + NTT(b) (mod q), and a sample of polynomials against the oracle."""
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = params.RNS_SETS["32k_16q"]
    L, num = 16, 1024
    ctx = nttb200.Context(n, qs, roots)
    psi, psiinv = _tables(oracle, n, qs, roots)
    g = torch.Generator(device="cuda").manual_seed(1234)
    qv = torch.tensor(qs, dtype=torch.int64, device="cuda").repeat(num // L).view(num, 1)
    a = torch.randint(0, 2**62, (num, n), dtype=torch.int64, device="cuda", generator=g) % qv
    b = torch.randint(0, 2**62, (num, n), dtype=torch.int64, device="cuda", generator=g) % qv
    s = (a + b) % qv
    fa, fb, fs = a.clone(), b.clone(), s.clone()
    for t in (fa, fb, fs):
        ctx.forward_ntt_batch(t, num, L)
    assert torch.equal((fa + fb) % qv, fs)
    for p in (0, 17, 1023):
        assert np.array_equal(to_host(fa[p]), oracle.forward_ntt_fast(to_host(a[p]), qs[p % L], psi[p % L]))
    ctx.inverse_ntt_batch(fa, num, L)
    assert torch.equal(fa, a)
    ctx.close()


def test_host_buffer_entry_point(oracle):
    import nttb200
    n, qs, roots = params.RNS_SETS["8k_3q"]
    psi, psiinv = _tables(oracle, n, qs, roots)
    num = 300   # several pipeline chunks, not a multiple of the chunk size
    ctx = nttb200.Context(n, qs, roots)
    a = np.concatenate([oracle.fill_uniform(n, qs[p % 3], 5 + p) for p in range(num)])
    out = np.empty_like(a)
    ctx.forward_ntt_batch_host(a, out, num, 3)
    for p in (0, 1, 149, 299):
        assert np.array_equal(out[p * n:(p + 1) * n], oracle.forward_ntt_fast(a[p * n:(p + 1) * n], qs[p % 3], psi[p % 3]))
    back = np.empty_like(a)
    ctx.inverse_ntt_batch_host(out, back, num, 3)
    assert np.array_equal(back, a)
    ctx.close()


def _pack_numpy(a, n, qbits, num, division):
    """The wire format restated in numpy: polynomial p as n * qbit bits, coefficient j at bit offset j * qbit, little-endian."""
    words = []
    for p in range(num):
        qb = qbits[p % division]
        bits = ((a[p * n:(p + 1) * n, None] >> np.arange(qb, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1)
        words.append(np.packbits(bits, bitorder="little").view(np.uint64))
    return np.concatenate(words)


@pytest.mark.parametrize("name,num", [("8k_3q", 300), ("16k_9q", 18)])
def test_packed_wire_format_and_host_packed_transforms(oracle, name, num):
    """SURVEY.md 8f-3 on the transform path: device pack / unpack against a numpy restatement of the layout, and the host-buffer
    transforms with both host arrays packed against the unpacked ones (several pipeline chunks, ragged last chunk)."""
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = params.RNS_SETS[name]
    L = len(qs)
    qbits = [int(q).bit_length() for q in qs]
    ctx = nttb200.Context(n, qs, roots)
    a = np.concatenate([oracle.fill_uniform(n, qs[p % L], 900 + p) for p in range(num)])
    a[0], a[1], a[2] = 0, 1, qs[0] - 1
    words = ctx.packed_words(num, L)
    assert words == sum(n // 64 * qbits[p % L] for p in range(num))
    d, pk = to_dev(a), to_dev(np.zeros(words, dtype=np.uint64))
    ctx.pack_polys(pk, d, num, L)
    want = _pack_numpy(a, n, qbits, num, L)
    assert np.array_equal(to_host(pk), want)
    back = to_dev(np.zeros_like(a))
    ctx.unpack_polys(back, pk, num, L)
    assert np.array_equal(to_host(back), a)
    # host-buffer transforms, packed in and out
    plain = np.empty_like(a)
    ctx.forward_ntt_batch_host(a, plain, num, L)
    pout = np.zeros(words, dtype=np.uint64)
    ctx.forward_ntt_batch_host_packed(want, pout, num, L)
    assert np.array_equal(pout, _pack_numpy(plain, n, qbits, num, L))
    pback = np.zeros(words, dtype=np.uint64)
    ctx.inverse_ntt_batch_host_packed(pout, pback, num, L)
    assert np.array_equal(pback, want)
    with pytest.raises(nttb200.NttB200Error):
        ctx.packed_words(num + 1, L)               # whole groups only
    ctx.close()


def test_invalid_arguments_fail_loudly():
    import nttb200
    with pytest.raises(nttb200.NttB200Error):
        nttb200.Context(1024, [12289], [7])          # n below 2^11
    with pytest.raises(nttb200.NttB200Error):
        nttb200.Context(2048, [137438691329], [5])   # not a primitive 2n-th root


@pytest.mark.parametrize("logn,limbs,num", [(16, 20, 40), (17, 16, 32)])
def test_c5_large_rings_many_limbs(oracle, logn, limbs, num):
    """BASELINE config 5: N = 2^16 .. 2^17 with 16+ RNS limbs (the reference stops at 32768 and at 16 limbs).  New 57-bit
    primes q = 1 mod 2N (nttb200.params.find_ntt_primes).  Oracle parity on a sample, round trip + linearity on everything."""
    import torch
    import nttb200
    from tests.gpu_util import to_host
    n = 1 << logn
    qs, roots = params.find_ntt_primes(57, n, limbs)
    ctx = nttb200.Context(n, qs, roots)
    g = torch.Generator(device="cuda").manual_seed(99)
    reps = -(-num // limbs)
    qv = torch.tensor(qs, dtype=torch.int64, device="cuda").repeat(reps)[:num].view(num, 1)
    a = torch.randint(0, 2**62, (num, n), dtype=torch.int64, device="cuda", generator=g) % qv
    b = torch.randint(0, 2**62, (num, n), dtype=torch.int64, device="cuda", generator=g) % qv
    fa, fb, fs = a.clone(), b.clone(), (a + b) % qv
    for t in (fa, fb, fs):
        ctx.forward_ntt_batch(t, num, limbs)
    assert torch.equal((fa + fb) % qv, fs)
    for p in (0, limbs - 1, num - 1):
        l = p % limbs
        psi, _ = oracle.fill_psi_tables(roots[l], qs[l], n)
        assert np.array_equal(to_host(fa[p]), oracle.forward_ntt_fast(to_host(a[p]), qs[l], psi))
    ctx.inverse_ntt_batch(fa, num, limbs)
    assert torch.equal(fa, a)
    ctx.close()


@pytest.mark.parametrize("logn,limbs,num", [(11, 1, 3), (12, 3, 4), (13, 1, 2), (14, 5, 6), (15, 16, 33), (16, 2, 3), (17, 1, 2)])
def test_fused_polynomial_product_vs_oracle(oracle, logn, limbs, num):
    """nttb200_poly_mul_batch (full_poly_mul_device, poly_arithmetic.cuh:296-310) against the schoolbook negacyclic product
    (helper.h:95-126, the check of 60bit_ntt_test.cu:85-98) at n = 2^11 and the oracle's NTT -> barrett -> INTT elsewhere."""
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = _ring(oracle, logn, limbs)
    psi, psiinv = _tables(oracle, n, qs, roots)
    ctx = nttb200.Context(n, qs, roots)
    a = np.concatenate([oracle.fill_uniform(n, qs[p % limbs], 0xA000 + p) for p in range(num)])
    b = np.concatenate([oracle.fill_uniform(n, qs[p % limbs], 0xB000 + p) for p in range(num)])
    a[0], a[1], a[2], b[0], b[1], b[2] = 0, 1, qs[0] - 1, qs[0] - 1, 0, qs[0] - 1
    da, db = to_dev(a), to_dev(b)
    ctx.poly_mul_batch(da, db, num, limbs)
    got = to_host(da)
    for p in range(0, num, max(1, num // 5)):
        q, l = qs[p % limbs], p % limbs
        ap, bp = a[p * n:(p + 1) * n], b[p * n:(p + 1) * n]
        if logn == 11:
            exp = oracle.ref_poly_mul(ap, bp, q)
        else:
            exp = oracle.inverse_ntt_fast(oracle.barrett(oracle.forward_ntt_fast(ap, q, psi[l]), oracle.forward_ntt_fast(bp, q, psi[l]), q), q, psiinv[l])
        assert np.array_equal(got[p * n:(p + 1) * n], exp), f"polynomial {p}"
    ctx.close()


def test_fused_product_full_size_equals_unfused_chain(oracle):
    """C2 size (1024 x 2^15, 16 limbs): the fused product equals forward_ntt_batch x2 -> coefficient-wise product -> inverse_ntt_batch
    done with separate calls, and ntt_domain_mul_inverse_batch equals the last two steps; commutativity a*b == b*a."""
    import torch
    import nttb200
    n, qs, roots = params.RNS_SETS["32k_16q"]
    L, num = 16, 1024
    ctx = nttb200.Context(n, qs, roots)
    g = torch.Generator(device="cuda").manual_seed(99)
    qv = torch.tensor(qs, dtype=torch.int64, device="cuda").repeat(num // L).view(num, 1)
    a = torch.randint(0, 2**62, (num, n), dtype=torch.int64, device="cuda", generator=g) % qv
    b = torch.randint(0, 2**62, (num, n), dtype=torch.int64, device="cuda", generator=g) % qv
    fa, fb = a.clone(), b.clone()
    ctx.forward_ntt_batch(fa, num, L)
    ctx.forward_ntt_batch(fb, num, L)
    prod = fa.clone()
    nttb200.barrett_batch(prod, fb, n, num, L, ctx.q_dev, ctx.mu_dev, ctx.qbit_dev)
    ref = prod.clone()
    ctx.inverse_ntt_batch(ref, num, L)
    x, y = a.clone(), b.clone()
    ctx.poly_mul_batch(x, y, num, L)
    assert torch.equal(x, ref)
    x2, y2 = b.clone(), a.clone()
    ctx.poly_mul_batch(x2, y2, num, L)
    assert torch.equal(x2, ref)
    z = fa.clone()
    ctx.ntt_domain_mul_inverse_batch(z, fb, num, L)
    assert torch.equal(z, ref)
    ctx.close()


@pytest.mark.parametrize("logn", [11, 15])
def test_final_reduction_arms_on_gpu(oracle, logn):
    """All three canonicalisation arms of the lazy forward transform in one batch (see tests/test_emu_ntt.py): a prime around
    3 * 2^53 (general 32-bit quotient arm), one just under 2^55 (pseudo-Mersenne arm), one below 2^32 (64-bit quotient arm)."""
    import nttb200
    from tests.gpu_util import to_dev, to_host
    n = 1 << logn
    q_mid, r_mid = params.find_ntt_primes(55, n, 1, below=3 << 53)
    q_pm, r_pm = params.find_ntt_primes(55, n, 1)
    q_small, r_small = params.find_ntt_primes(30, n, 1)
    qs, roots = q_mid + q_pm + q_small, r_mid + r_pm + r_small
    psi, psiinv = _tables(oracle, n, qs, roots)
    num = 9
    a = np.concatenate([oracle.fill_uniform(n, qs[i % 3], 0xF1A7 + i) for i in range(num)])
    ctx = nttb200.Context(n, qs, roots)
    d = to_dev(a)
    ctx.forward_ntt_batch(d, num, 3)
    assert np.array_equal(to_host(d), _expect(oracle, a, n, qs, psi, psiinv, num, 3, False))
    ctx.inverse_ntt_batch(d, num, 3)
    assert np.array_equal(to_host(d), a)
    b = np.concatenate([oracle.fill_uniform(n, qs[i % 3], 0xBEE5 + i) for i in range(num)])
    da, db = to_dev(a), to_dev(b)
    ctx.poly_mul_batch(da, db, num, 3)
    got = to_host(da)
    for i in (0, 1, 2):
        q, l = qs[i % 3], i % 3
        exp = oracle.inverse_ntt_fast(oracle.barrett(oracle.forward_ntt_fast(a[i * n:(i + 1) * n], q, psi[l]),
                                                     oracle.forward_ntt_fast(b[i * n:(i + 1) * n], q, psi[l]), q), q, psiinv[l])
        assert np.array_equal(got[i * n:(i + 1) * n], exp)
    ctx.close()
