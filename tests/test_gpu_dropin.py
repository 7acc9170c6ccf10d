"""The reference's OWN driver programs, unmodified, compiled against include/dropin/ + libnttb200.so (oracle/ref/Makefile),
and GPU-vs-GPU parity of the library against the reference's kernels rebuilt for sm_100a (oracle/_ref/ref_dump).
The binaries are built in the build container (where /root/reference exists) and travel to the GPU box."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from nttb200 import params  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _run(name, *args, timeout=600):
    exe = os.path.join(REF, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_decryption_test_drops_in():
    out = _run("dropin_decryption_test")
    assert "Computations are correct." in out and "[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, ]" in out


def test_reference_demo_drops_in():
    """demo.cu as committed: n = 32768, 16 primes, keygen -> encrypt -> decrypt of a random message."""
    out = _run("dropin_demo")
    assert "Computations are correct." in out and "# of qs: 16" in out


def test_reference_ntt_test_drops_in_with_check_enabled():
    out = _run("dropin_60bit_ntt_test_check")
    assert "error" not in out and "n = 2048" in out


def test_reference_keygen_test_histogram_identical():
    """keygen_test.cu prints the ternary histogram of 341 M keystream bytes: same counts from both builds."""
    a = _run("dropin_keygen_test")
    b = _run("orig_keygen_test")
    assert a == b and "Number of -1 generated" in a


@pytest.mark.parametrize("name,num", [("32k_16q", 48), ("8k_3q", 7), ("4k_3q", 6)])
def test_ntt_gpu_vs_reference_kernels(name, num):
    """Raw forward / inverse NTT outputs: library (context, Shoup) and stateless (Barrett) vs the reference's kernels."""
    import nttb200
    from oracle import oracle as orc
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = params.RNS_SETS[name]
    r = len(qs)
    with tempfile.TemporaryDirectory() as d:
        _run("ref_dump", "ntt", name, str(num), d)
        ref_fwd = np.fromfile(os.path.join(d, "ref_fwd.bin"), dtype=np.uint64)
        ref_inv = np.fromfile(os.path.join(d, "ref_inv.bin"), dtype=np.uint64)
    a = np.concatenate([orc.fill_uniform(n, qs[p % r], 0x5EED0000 + p) for p in range(num)])
    assert np.array_equal(ref_inv, a)                      # the reference round-trips these inputs
    ctx = nttb200.Context(n, qs, roots)
    dv = to_dev(a)
    ctx.forward_ntt_batch(dv, num, r)
    assert np.array_equal(to_host(dv), ref_fwd)
    ctx.inverse_ntt_batch(dv, num, r)
    assert np.array_equal(to_host(dv), ref_inv)
    nttb200.forwardNTT_batch(dv, n, ctx.psi_table, num, r, ctx.q_dev, ctx.mu_dev, ctx.qbit_dev)
    assert np.array_equal(to_host(dv), ref_fwd)
    nttb200.inverseNTT_batch(dv, n, ctx.psiinv_table, num, r, ctx.q_dev, ctx.mu_dev, ctx.qbit_dev)
    assert np.array_equal(to_host(dv), ref_inv)
    ctx.close()


@pytest.mark.parametrize("name", ["4k_3q", "8k_4q", "16k_5q", "32k_9q", "32k_16q"])
def test_bfv_gpu_vs_reference_pipelines(name):
    """keygen_rns / encryption_rns / decryption_rns of the reference (rebuilt) vs the library on the same seeds: keystream,
    secret key, public key, gaussian draws (normcdfinvf: pinned here, GPU vs GPU), ciphertext incl. padding, plaintext."""
    import torch
    import nttb200
    from oracle import oracle as orc
    from tests.gpu_util import to_dev, to_host
    n, qs, roots = params.RNS_SETS[name]
    r = len(qs)
    rn = r * n
    with tempfile.TemporaryDirectory() as d:
        out = _run("ref_dump", "bfv", name, d)
        assert "roundtrip ok" in out
        ref = {k: np.fromfile(os.path.join(d, f"ref_{k}.bin"), dtype=np.uint8) for k in ("keygen_in", "sk", "pk", "temp", "c", "e", "plain")}
    u64v = lambda k: ref[k].view(np.uint64)
    bfv = nttb200.Bfv(n, qs, roots)
    ctx_psi = nttb200.Context(n, qs, roots)
    R = orc.Ring(n, qs, roots)
    dq, dmu, dqb = ctx_psi.q_dev, ctx_psi.mu_dev, ctx_psi.qbit_dev
    inb = torch.zeros(9 * rn + 4 * n, dtype=torch.uint8, device="cuda")
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    temp = torch.zeros(rn, dtype=torch.int64, device="cuda")
    nttb200.keygen_rns(inb, r, n, sk, pk, temp, ctx_psi.psi_table, ctx_psi.psiinv_table, dq, dmu, dqb)
    assert np.array_equal(to_host(inb), ref["keygen_in"])
    assert np.array_equal(to_host(sk), u64v("sk"))
    assert np.array_equal(to_host(pk), u64v("pk"))
    # gaussian draws: the reference stores residues per limb, the library n signed draws
    es = to_host(temp).view(np.int32)[:n].astype(np.int64)
    for l in range(r):
        assert np.array_equal(np.where(es < 0, es + qs[l], es).astype(np.uint64), u64v("temp")[l * n:(l + 1) * n])
    m = orc.fill_uniform(n, params.T, 0xC0FFEE)
    c = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    e = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    nttb200.encryption_rns(c, pk, inb, e, n, ctx_psi.psi_table, ctx_psi.psiinv_table, to_dev(m), to_dev(R.qi_div_t), params.T, r, dq, dmu, dqb,
                           to_dev(R.inv_q_last_mod_q))
    assert np.array_equal(to_host(c), u64v("c"))
    # the batched context API produces the same key pair and ciphertext for nonce 0
    sk2 = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk2 = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk2, pk2)
    assert np.array_equal(to_host(sk2), u64v("sk")) and np.array_equal(to_host(pk2), u64v("pk"))
    c2 = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.encrypt(c2, pk2, to_dev(m))
    assert np.array_equal(to_host(c2), u64v("c"))
    outp = torch.zeros(n, dtype=torch.int64, device="cuda")
    bfv.decrypt(outp, c2, sk2)
    assert np.array_equal(to_host(outp), u64v("plain")) and np.array_equal(to_host(outp), m)
    bfv.close()
    ctx_psi.close()


def test_native_cpp_example_runs():
    """examples/bfv_batched.cpp: the C ABI used from plain C++ (keygen, load_keys, batched encrypt, wire format, homomorphic add, decrypt)."""
    exe = os.path.join(ROOT, "ntt-cuda_b200", "build", "example_bfv_batched")
    if not os.path.exists(exe):
        pytest.skip("example not built (python -c 'import __graft_entry__ as g; g.build()')")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
