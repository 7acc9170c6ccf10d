"""world_size-2 gloo tests (CPU) of the N > 1 host logic: sharding arithmetic and the limb-sharded decryption control flow
(partial sums -> ONE all-reduce -> finish), with the emulator build of the kernels standing in for the GPU library."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_arithmetic():
    from nttb200.distributed import shard_batch, shard_limbs
    for total in (1, 7, 8, 1024, 4097):
        for world in (1, 2, 4, 8):
            parts = [shard_batch(total, world, r) for r in range(world)]
            assert sum(c for _, c in parts) == total
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    for rp in (2, 3, 8, 15):
        for world in (1, 2, 4, 8):
            parts = [shard_limbs(rp, world, r) for r in range(world)]
            assert sum(c for _, c in parts) == rp and parts[0][0] == 0
            assert all(f + c <= rp for f, c in parts)


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
        import torch
        import torch.distributed as dist
        from nttb200 import params
        from nttb200.distributed import ciphertext_limb_shard, decrypt_limb_sharded, shard_limbs
        from oracle import oracle as orc
        from tests import emu
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        n, qs, roots = params.RNS_SETS["8k_4q"]
        R = orc.Ring(n, qs, roots)
        r, rp = R.r, R.r - 1
        B = 2
        sk, pk, _, _ = orc.keygen_rns(R)
        ms, cs = [], []
        for k in range(B):
            orc.set_nonce(k)
            m = orc.fill_uniform(n, R.t, 900 + k)
            c, _ = orc.encryption_rns(R, pk, m)
            ms.append(m)
            cs.append(c)
        orc.set_nonce(0)
        c_all = np.concatenate(cs)
        first, count = shard_limbs(rp, world, rank)
        c_shard = ciphertext_limb_shard(c_all, n, r, first, count, batch=B)
        sk_shard = np.ascontiguousarray(sk.reshape(r, n)[first:first + count]).reshape(-1)
        bfv = emu.EmuBfv(emu.EmuRing(R))

        def all_reduce_sum(buf):
            t = torch.from_numpy(buf.view(np.int64))      # shares memory: in-place sum lands in `buf`
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

        out = decrypt_limb_sharded(bfv, c_shard, sk_shard, first, count, B, all_reduce_sum, lambda cnt: np.zeros(cnt, dtype=np.uint64))
        ok = all(np.array_equal(out[k * n:(k + 1) * n], ms[k]) for k in range(B))
        # limb-sharded ENCRYPTION on the sub-ring {owned limbs, last limb}: no communication, same bits as the full ciphertext
        from nttb200.distributed import encrypt_limb_sharded, public_key_limb_shard, sub_ring
        if count > 0:
            q_sub, roots_sub = sub_ring(qs, roots, first, count)
            sub = emu.EmuBfv(emu.EmuRing(orc.Ring(n, q_sub, roots_sub)))
            c_enc = np.zeros(B * 2 * (count + 1) * n, dtype=np.uint64)
            # item k of the full run used nonce k: encrypt item by item with the same nonces
            for k in range(B):
                one = np.zeros(2 * (count + 1) * n, dtype=np.uint64)
                encrypt_limb_sharded(sub, one, public_key_limb_shard(pk, n, r, first, count), ms[k], 1, nonce0=k)
                c_enc[k * one.size:(k + 1) * one.size] = one
            full = c_all.reshape(B, 2, r, n)
            mine = c_enc.reshape(B, 2, count + 1, n)
            ok = ok and np.array_equal(mine[:, :, :count, :], full[:, :, first:first + count, :])
            # ... and that shard feeds limb-sharded decryption directly (shard_half_limbs = count + 1)
            out2 = decrypt_limb_sharded(bfv, c_enc, sk_shard, first, count, B, all_reduce_sum, lambda cnt: np.zeros(cnt, dtype=np.uint64),
                                        shard_half_limbs=count + 1)
        else:
            out2 = decrypt_limb_sharded(bfv, None, None, first, count, B, all_reduce_sum, lambda cnt: np.zeros(cnt, dtype=np.uint64))
        ok = ok and all(np.array_equal(out2[k * n:(k + 1) * n], ms[k]) for k in range(B))
        # and identical to the single-device oracle decryption
        for k in range(B):
            plain, _ = orc.decryption_rns(R, cs[k], sk)
            ok = ok and np.array_equal(out[k * n:(k + 1) * n], plain)
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 3])
def test_limb_sharded_decrypt_gloo(oracle, world):
    import torch.multiprocessing as mp
    from tests import emu
    emu.lib()                 # build the emulator once, before forking
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=600) for _ in procs]
    for p_ in procs:
        p_.join(timeout=60)
    for rank, ok, err in res:
        assert ok, f"rank {rank}: {err}"
