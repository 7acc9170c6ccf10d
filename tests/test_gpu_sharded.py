"""GPU parity (one GPU) of round 2's BFV paths through the C ABI:
  * fused-epilogue encryption (5 launches, epilogue in the store of the last inverse kernel) == the separate-epilogue path == oracle,
  * nttb200_bfv_encrypt_sharded / _decrypt_sharded at world size 1 (same code path the multi-GPU runs take per rank),
  * the limb-sharded decryption split: every virtual rank of a world of G in {2, 3, 8} runs its plan tiles through
    nttb200_bfv_decrypt_partial_tile (packed and unpacked), the partial sums are SUM-reduced on the device as NCCL would, the owner
    rounds with nttb200_bfv_decrypt_finish_tile -- == nttb200_bfv_decrypt == oracle, and the reference's KAT (decryption_test.cu:348,355),
  * the round-1 entry points nttb200_bfv_decrypt_partial / _finish split into G shards.
The real multi-rank run (NCCL) is scripts/multigpu_check.py under torchrun; bench.py --gpus N re-checks bit-identity every run."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from nttb200 import params  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "decryption_kat.npz")


def _setup(oracle, name, B, seed=0x5A):
    import torch
    import nttb200
    n, q, roots = params.RNS_SETS[name]
    R = oracle.Ring(n, q, roots)
    bfv = nttb200.Bfv(n, q, roots)
    rn = R.r * n
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk)
    bfv.load_keys(sk, pk)
    m = np.concatenate([oracle.fill_uniform(n, R.t, seed + k) for k in range(B)])
    m[:4] = [0, R.t - 1, 1, R.t - 1]
    return R, bfv, sk, pk, m


@pytest.mark.parametrize("name", ["4k_3q", "8k_4q", "16k_5q", "16k_9q", "32k_16q"])
def test_fused_epilogue_encryption(oracle, name):
    import torch
    import nttb200  # noqa: F401
    from tests.gpu_util import to_dev, to_host
    B = 3
    R, bfv, sk, pk, m = _setup(oracle, name, B)
    n, rn = R.n, R.r * R.n
    md = to_dev(m)
    c_new = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    c_old = torch.zeros_like(c_new)
    bfv.set_fused_epilogue(True)
    bfv.encrypt(c_new, None, md, batch=B, nonce0=21)             # loaded key, epilogue in the store of the last inverse kernel (A/B variant)
    bfv.set_fused_epilogue(False)
    bfv.encrypt(c_old, None, md, batch=B, nonce0=21)             # loaded key, separate epilogue kernels (default)
    assert torch.equal(c_new, c_old), "fused-epilogue ciphertext (padding limb included) != separate-epilogue ciphertext"
    c_pk = torch.zeros_like(c_new)
    bfv.encrypt(c_pk, pk, md, batch=B, nonce0=21)                # explicit key: unfused kernels
    assert torch.equal(c_new, c_pk)
    out = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    bfv.decrypt(out, c_new.clone(), None, batch=B)
    assert np.array_equal(to_host(out), m)
    # the oracle decrypts what the GPU encrypted
    hc, hsk = to_host(c_new), to_host(sk)
    for k in (0, B - 1):
        plain, _ = oracle.decryption_rns(R, hc[k * 2 * rn:(k + 1) * 2 * rn], hsk)
        assert np.array_equal(plain, m[k * n:(k + 1) * n])
    bfv.close()


@pytest.mark.parametrize("name", ["8k_4q", "32k_16q"])
def test_sharded_calls_world1(oracle, name):
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    B = 4
    R, bfv, sk, pk, m = _setup(oracle, name, B, seed=0x77)
    n, r = R.n, R.r
    rn = r * n
    md = to_dev(m)
    comm = nttb200.Comm.single()
    words = bfv.shard_words(comm, B)
    assert words == B * 2 * (r - 1) * n
    shard = torch.zeros(words, dtype=torch.int64, device="cuda")
    bfv.encrypt_sharded(comm, shard, md, B, nonce0=5)
    full = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    bfv.encrypt(full, None, md, batch=B, nonce0=5)
    assert torch.equal(shard.view(B, 2, r - 1, n), full.view(B, 2, r, n)[:, :, :r - 1, :])
    bfv.set_fused_epilogue(True)                                 # the A/B variant of the sharded call's last inverse kernel
    shard2 = torch.zeros_like(shard)
    bfv.encrypt_sharded(comm, shard2, md, B, nonce0=5)
    bfv.set_fused_epilogue(False)
    assert torch.equal(shard2, shard)
    back = torch.zeros_like(full)
    bfv.shard_to_full(1, 0, back, shard, B)
    assert torch.equal(back.view(B, 2, r, n)[:, :, :r - 1, :], full.view(B, 2, r, n)[:, :, :r - 1, :])
    again = torch.zeros_like(shard)
    bfv.shard_from_full(1, 0, again, full, B)
    assert torch.equal(again, shard)
    out = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    bfv.decrypt_sharded(comm, out, shard, B)
    assert np.array_equal(to_host(out), m)
    comm.close()
    bfv.close()


def _virtual_world_decrypt(bfv, nttb200, torch, c_full, B, G, packed, legacy=False):
    """All G virtual ranks on one device: tiles of the product's plan -> partial sums -> device-side SUM (what ncclReduce does) -> rounding."""
    n, r = bfv.n, bfv.r
    pw = n + n // 4 if packed else 2 * n
    per = B // G
    sums = torch.zeros(B * pw, dtype=torch.int64, device="cuda")
    for g in range(G):
        plan, words = nttb200.shard_plan(r - 1, n, B, G, g)
        shard = torch.zeros(max(words, 1), dtype=torch.int64, device="cuda")
        bfv.shard_from_full(G, g, shard, c_full, B)
        for (it, items, f, cnt, off) in plan:
            if not cnt:
                continue
            part = torch.zeros(items * pw, dtype=torch.int64, device="cuda")
            if legacy:
                sk_shard = bfv._sk_full.view(r, n)[f:f + cnt].contiguous().view(-1)
                bfv.decrypt_partial(part, shard[off:], sk_shard, f, cnt, batch=items)
            else:
                bfv.decrypt_partial_tile(part, packed, shard[off:], f, cnt, batch=items)
            sums[it * pw:(it + items) * pw] += part            # int64 wrap-around == the u64 sum
    out = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    if legacy:
        bfv.decrypt_finish(out, sums, batch=B)
        return out
    for g in range(G):                                          # every owner rounds its own block
        sl = slice(g * per * n, (g + 1) * per * n)
        if bfv.t <= 65536:
            p16 = torch.zeros(per * n, dtype=torch.int16, device="cuda")
            bfv.decrypt_finish_tile(p16, True, sums[g * per * pw:], packed, batch=per)
            out[sl] = p16.to(torch.int64) & 0xFFFF
        else:
            bfv.decrypt_finish_tile(out[sl], False, sums[g * per * pw:], packed, batch=per)
    return out


@pytest.mark.parametrize("name,G", [("8k_4q", 2), ("8k_4q", 3), ("32k_16q", 2), ("32k_16q", 3), ("32k_16q", 8)])
def test_limb_sharded_decryption_virtual_ranks(oracle, name, G):
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    B = G * 2 if G < 8 else 8
    R, bfv, sk, pk, m = _setup(oracle, name, B, seed=0x99)
    bfv._sk_full = sk
    n, rn = R.n, R.r * R.n
    c = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    bfv.encrypt(c, None, to_dev(m), batch=B, nonce0=1)
    ref = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    bfv.decrypt(ref, c.clone(), None, batch=B)
    assert np.array_equal(to_host(ref), m)
    for packed in (True, False):
        out = _virtual_world_decrypt(bfv, nttb200, torch, c, B, G, packed)
        assert torch.equal(out, ref), f"G={G} packed={packed}"
    out = _virtual_world_decrypt(bfv, nttb200, torch, c, B, G, False, legacy=True)      # round-1 entries nttb200_bfv_decrypt_partial / _finish
    assert torch.equal(out, ref)
    hc, hsk = to_host(c), to_host(sk)
    plain, _ = oracle.decryption_rns(R, hc[:2 * rn], hsk)
    assert np.array_equal(to_host(ref)[:n], plain)
    bfv.close()


def test_limb_sharded_decryption_kat(oracle):
    """The reference's only golden vector through the sharded split: c_host / sk_host -> i % 10 (decryption_test.cu:348,355)."""
    import torch
    import nttb200
    from tests.gpu_util import to_dev, to_host
    g = np.load(GOLD)
    n, q, roots = params.RNS_SETS["4k_3q"]
    bfv = nttb200.Bfv(n, q, roots)
    r = len(q)
    sk = np.zeros(r * n, dtype=np.uint64)
    sk[:2 * n] = g["sk_host"]
    skd = to_dev(sk)
    bfv.load_keys(skd, None)
    bfv._sk_full = skd
    B, G = 2, 2
    c = to_dev(np.tile(g["c_host"], B))
    expect = np.tile(np.arange(n, dtype=np.uint64) % 10, B)
    for packed in (True, False):
        out = _virtual_world_decrypt(bfv, nttb200, torch, c, B, G, packed)
        assert np.array_equal(to_host(out), expect)
    out = _virtual_world_decrypt(bfv, nttb200, torch, c, B, G, False, legacy=True)
    assert np.array_equal(to_host(out), expect)
    comm = nttb200.Comm.single()
    shard = torch.zeros(bfv.shard_words(comm, B), dtype=torch.int64, device="cuda")
    bfv.shard_from_full(1, 0, shard, c, B)
    o2 = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    bfv.decrypt_sharded(comm, o2, shard, B)
    assert np.array_equal(to_host(o2), expect)
    comm.close()
    bfv.close()


def test_bfv_create_rejects_unsupported_parameters():
    """t that is not a power of two <= 2^32, q_i != 1 mod t, gamma != 1 mod t, composite / oversized gamma -> EINVAL."""
    import nttb200
    n, q, roots = params.RNS_SETS["8k_4q"]
    for kw in (dict(t=1000), dict(t=1 << 15), dict(t=4096), dict(gamma=params.GAMMA + 2), dict(gamma=(1 << 62) + 1025), dict(t=1 << 33)):
        with pytest.raises(nttb200.NttB200Error):
            nttb200.Bfv(n, q, roots, **kw)
    nttb200.Bfv(n, q, roots, t=2048).close()        # gamma = 1 mod 2^11 and q_i = 1 mod 2^14: valid


def test_host_buffer_bfv_and_wire_formats(oracle):
    """nttb200_bfv_encrypt_host / _decrypt_host (packed and reference layout, more than one chunk), key wire format, host pack/unpack."""
    import torch
    import nttb200  # noqa: F401
    from tests.gpu_util import to_dev, to_host
    B = 40                                       # 32k_16q: 16 items per 128 MiB chunk -> three chunks
    R, bfv, sk, pk, m = _setup(oracle, "32k_16q", B, seed=0x31)
    n, r = R.n, R.r
    rn = r * n
    pw = bfv.packed_words()
    mh = torch.from_numpy(m.view(np.int64)).pin_memory()
    ref = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    bfv.encrypt(ref, None, to_dev(m), batch=B, nonce0=9)
    for packed in (True, False):
        ch = torch.zeros(B * (pw if packed else 2 * rn), dtype=torch.int64).pin_memory()
        bfv.encrypt_host(ch, mh, B, nonce0=9, packed=packed)
        if packed:
            pk_dev = torch.zeros(B * pw, dtype=torch.int64, device="cuda")
            bfv.pack(pk_dev, ref, batch=B)
            assert torch.equal(ch, pk_dev.cpu())
            h2 = torch.zeros(B * pw, dtype=torch.int64)
            bfv.pack_host(h2, ref, batch=B)
            assert torch.equal(h2, ch)
            back = torch.full((B * 2 * rn,), -1, dtype=torch.int64, device="cuda")
            bfv.unpack_host(back, h2, batch=B)
            assert torch.equal(back.view(B, 2, r, n)[:, :, :r - 1], ref.view(B, 2, r, n)[:, :, :r - 1])
        else:
            assert torch.equal(ch, ref.cpu())
        out = torch.zeros(B * n, dtype=torch.int64).pin_memory()
        bfv.decrypt_host(out, ch, B, packed=packed)
        assert np.array_equal(out.numpy().view(np.uint64), m)
    # keys: all r limbs, n * qbit_l bits each
    for key, polys in ((sk, r), (pk, 2 * r)):
        words = bfv.key_packed_words(polys)
        assert words == polys // r * sum(n // 64 * int(b) for b in R.qbit)
        p = torch.zeros(words, dtype=torch.int64, device="cuda")
        bfv.pack_key(p, key, polys)
        back = torch.zeros_like(key)
        bfv.unpack_key(back, p, polys)
        assert torch.equal(back, key)
        # format check against numpy on the first limb
        hk = to_host(key)[:n]
        qb = int(R.qbit[0])
        bits = ((hk[:, None] >> np.arange(qb, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1)
        assert np.array_equal(to_host(p)[:n // 64 * qb], np.packbits(bits, bitorder="little").view(np.uint64))
    bfv.close()
