"""CPU tests of round 2's pipelines on the emulator: the fused-epilogue encryption (epilogue in the store of the last inverse kernel),
the fused / packed sharded-decryption building blocks, the product's partition (nttb200_shard_plan) and the schedule of
csrc/sharded.cu at world sizes 1, 2 and 3 (gloo), all bit-exact against the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

from nttb200 import params
from tests import emu, sharded_sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_ciphertexts(oracle, R, pk, m, es, batch, nonce0):
    n = R.n
    out = []
    for k in range(batch):
        oracle.set_nonce(nonce0 + k)
        e = es[k * 2 * n:(k + 1) * 2 * n].astype(np.int32)
        oc, _ = oracle.encryption_rns(R, pk, m[k * n:(k + 1) * n], e0_samples=np.ascontiguousarray(e[:n]), e1_samples=np.ascontiguousarray(e[n:]))
        out.append(oc)
    oracle.set_nonce(0)
    return np.concatenate(out)


def test_shard_plan_partition():
    """Every (limb, block) tile is owned exactly once, every rank owns rp tiles, a rank's limbs of one block are contiguous."""
    for rp in (2, 3, 4, 8, 15, 16, 31):
        for world in (1, 2, 3, 4, 8):
            batch = 4 * world
            seen = np.zeros((rp, world), dtype=np.int32)
            for g in range(world):
                plan, words = sharded_sim.shard_plan(rp, 2048, batch, world, g)
                tiles, off = 0, 0
                for j, (it, items, f, cnt, o) in enumerate(plan):
                    assert (it, items) == (j * 4, 4) and o == off
                    seen[f:f + cnt, j] += 1
                    tiles += cnt
                    off += items * 2 * cnt * 2048
                assert tiles == rp and words == off
            assert (seen == 1).all(), (rp, world)


@pytest.mark.parametrize("fused", [False, True], ids=["epilogue_kernel", "epilogue_in_store"])
def test_fused_epilogue_encryption_matches_oracle(oracle, fused):
    """run_encrypt_v2's launch sequence (csrc/bfv.cu) on the emulator: sampling as signed bytes, strided forward pass generating u,
    fused contig kernel, strided inverse of the dropped limb with `+ e` / rounding in its store, strided inverse of the other limbs
    with mod-switch + Delta*m in its store -- ciphertext INCLUDING the padding limb == oracle, extremes of m included."""
    n, q, roots = params.RNS_SETS["8k_4q"]
    R = oracle.Ring(n, q, roots)
    er = emu.EmuRing(R)
    r = R.r
    sk, pk, _ = emu.bfv(0, er, 0)
    B, nonce0 = 2, 5
    m = np.concatenate([oracle.fill_uniform(n, R.t, 0xE1 + k) for k in range(B)])
    m[:4] = [0, R.t - 1, 1, R.t - 1]
    blk = emu.EmuBlocks(er, sk=sk, pk=pk)
    blk.fused = fused
    ub = np.zeros(B * n, dtype=np.uint8)
    es = np.zeros(B * 2 * n, dtype=np.int8)
    c = np.zeros(B * 2 * r * n, dtype=np.uint64)
    blk.enc_sample(ub, es, B, nonce0, 1, 1)
    assert np.array_equal(ub[:n], oracle.salsa20_keystream(n, b"\x01" * 32, nonce0)) and np.abs(es).max() <= 19
    blk.enc_front(c, r, 0, r, B, ub)
    cl = c[(r - 1) * n:]
    blk.enc_finish_last(cl, 2 * r * n, r * n, es, B)
    blk.enc_finish_limbs(c, r, 0, r - 1, B, cl, 2 * r * n, r * n, es, m)
    assert np.array_equal(c, _oracle_ciphertexts(oracle, R, pk, m, es, B, nonce0))
    # ... and the fused / packed decryption blocks give the messages back (all limbs in one window, world size 1)
    out = sharded_sim.decrypt_sharded(blk, sharded_sim.NoColl(), _drop_padding(c, n, r, B), B)
    assert np.array_equal(out, m)


def _drop_padding(c, n, r, B):
    return np.ascontiguousarray(c.reshape(B, 2, r, n)[:, :, :r - 1, :]).reshape(-1)


def test_sharded_schedule_world1_matches_oracle(oracle):
    n, q, roots = params.RNS_SETS["4k_3q"]
    R = oracle.Ring(n, q, roots)
    er = emu.EmuRing(R)
    sk, pk, _ = emu.bfv(0, er, 0)
    B = 2
    m = np.concatenate([oracle.fill_uniform(n, R.t, 0x51 + k) for k in range(B)])
    blk = emu.EmuBlocks(er, sk=sk, pk=pk)
    shard, es = sharded_sim.encrypt_sharded(blk, sharded_sim.NoColl(), m, B, 9)
    oc = _oracle_ciphertexts(oracle, R, pk, m, es, B, 9)
    assert np.array_equal(shard, _drop_padding(oc, n, R.r, B))
    assert np.array_equal(sharded_sim.decrypt_sharded(blk, sharded_sim.NoColl(), shard, B), m)


@pytest.mark.parametrize("tbits", [16, 17])
def test_unpacked_partials_when_t_is_large(oracle, tbits):
    """rp * (t - 1) >= 2^16: the partial sums stay two u64 per coefficient; t > 2^16: the plaintext is gathered as u64."""
    n = 2048
    qs, _ = params.find_ntt_primes(50, 65536, 3)              # q = 1 mod 2^17
    roots = []
    for q in qs:
        g = 2
        while pow(pow(g, (q - 1) // (2 * n), q), n, q) != q - 1:
            g += 1
        roots.append(pow(g, (q - 1) // (2 * n), q))
    gamma = 2305843009211596801                              # prime, = 1 mod 2^17 (the reference's gamma is 1 only mod 2^11: t <= 2048)
    R = oracle.Ring(n, qs, roots, t=1 << tbits, gamma=gamma, gamma_bits=61)
    er = emu.EmuRing(R)
    sk, pk, _ = emu.bfv(0, er, 0)
    m = oracle.fill_uniform(n, R.t, 0x77)
    m[:2] = [0, R.t - 1]
    blk = emu.EmuBlocks(er, sk=sk, pk=pk)
    shard, es = sharded_sim.encrypt_sharded(blk, sharded_sim.NoColl(), m, 1, 0)
    oc = _oracle_ciphertexts(oracle, R, pk, m, es, 1, 0)
    assert np.array_equal(shard, _drop_padding(oc, n, R.r, 1))
    assert np.array_equal(sharded_sim.decrypt_sharded(blk, sharded_sim.NoColl(), shard, 1), m)


@pytest.mark.parametrize("fused", [False, True], ids=["epilogue_kernel", "epilogue_in_store"])
def test_fused_epilogue_with_a_much_larger_dropped_limb(oracle, fused):
    """q_last > 2 q_i (mixed-size sets such as 16k_9q: 48 / 49 / 50-bit primes): the fused epilogue reduces c_last mod q_i itself."""
    n = 2048
    qa, ra = params.find_ntt_primes(48, n, 2)
    qb, rb = params.find_ntt_primes(51, n, 1)
    qs, roots = qa + qb, ra + rb
    assert qs[-1] > 2 * qs[0]
    R = oracle.Ring(n, qs, roots)
    er = emu.EmuRing(R)
    sk, pk, _ = emu.bfv(0, er, 0)
    B = 2
    m = np.concatenate([oracle.fill_uniform(n, R.t, 0x3C + k) for k in range(B)])
    blk = emu.EmuBlocks(er, sk=sk, pk=pk)
    blk.fused = fused
    shard, es = sharded_sim.encrypt_sharded(blk, sharded_sim.NoColl(), m, B, 0)
    oc = _oracle_ciphertexts(oracle, R, pk, m, es, B, 0)
    assert np.array_equal(shard, _drop_padding(oc, n, R.r, B))
    assert np.array_equal(sharded_sim.decrypt_sharded(blk, sharded_sim.NoColl(), shard, B), m)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        for pth in (ROOT, os.path.join(ROOT, "ntt-cuda_b200")):
            if pth not in sys.path:
                sys.path.insert(0, pth)
        import torch.distributed as dist
        from oracle import oracle as orc
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        n, qs, roots = params.RNS_SETS["8k_4q"]
        R = orc.Ring(n, qs, roots)
        er = emu.EmuRing(R)
        r = R.r
        B = 2 * world                                 # two items per block: two pieces per block in the decryption
        sk, pk, _, _ = orc.keygen_rns(R)
        m = np.concatenate([orc.fill_uniform(n, R.t, 0x900 + k) for k in range(B)])
        blk = emu.EmuBlocks(er, sk=sk, pk=pk)
        coll = sharded_sim.GlooColl(dist, world, rank)
        shard, es = sharded_sim.encrypt_sharded(blk, coll, m, B, 3)
        # every rank now holds every item's draws: the oracle's ciphertexts, restricted to this rank's tiles
        oc = _oracle_ciphertexts(orc, R, pk, m, es, B, 3).reshape(B, 2, r, n)
        plan, words = sharded_sim.shard_plan(r - 1, n, B, world, rank)
        ok = True
        for (it, items, f, cnt, off) in plan:
            if cnt:
                ok = ok and np.array_equal(shard[off:off + items * 2 * cnt * n].reshape(items, 2, cnt, n), oc[it:it + items, :, f:f + cnt, :])
        out = sharded_sim.decrypt_sharded(blk, coll, shard, B)
        ok = ok and np.array_equal(out, m)
        for k in range(B):
            plain, _ = orc.decryption_rns(R, oc[k].reshape(-1), sk)
            ok = ok and np.array_equal(out[k * n:(k + 1) * n], plain)
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_schedule_gloo(oracle, world):
    """The N > 1 control flow of csrc/sharded.cu (partition, all-gather of the finished dropped limb and the draws, per-block reduce of
    packed partial sums to the owner, rounding on the owner, all-gather of 16-bit plaintexts) over gloo, emulator kernels."""
    import torch.multiprocessing as mp
    emu.lib()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r_, world, port, q)) for r_ in range(world)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=900) for _ in procs]
    for p_ in procs:
        p_.join(timeout=60)
    for rank, ok, err in res:
        assert ok, f"rank {rank}: {err}"
