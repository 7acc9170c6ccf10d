"""Test infrastructure: the schedule of csrc/sharded.cu (nttb200_bfv_encrypt_sharded / _decrypt_sharded) restated in Python over the
emulator's building blocks, with the collectives supplied by the caller (torch.distributed over gloo in the CPU tests, or no-ops for
world size 1).  The partition itself comes from the PRODUCT's host-only nttb200_shard_plan (libnttb200.so loads without a GPU)."""
import ctypes as C

import numpy as np


class ShardBlock(C.Structure):
    _fields_ = [("first_item", C.c_uint), ("items", C.c_uint), ("first_limb", C.c_uint), ("limb_count", C.c_uint), ("offset", C.c_size_t)]


def shard_plan(rp, n, batch, world, rank):
    """-> (list of (first_item, items, first_limb, limb_count, offset), shard_words) from the product library."""
    import nttb200
    lib = nttb200.lib()
    blocks = (ShardBlock * world)()
    words = C.c_size_t(0)
    rc = lib.nttb200_shard_plan(C.c_uint(rp), C.c_uint(n), C.c_uint(batch), C.c_uint(world), C.c_uint(rank), blocks, C.byref(words))
    assert rc == 0, rc
    return [(b.first_item, b.items, b.first_limb, b.limb_count, b.offset) for b in blocks], words.value


class NoColl:
    """world size 1"""
    world, rank = 1, 0

    def all_gather_inplace(self, buf, per):
        pass

    def reduce_sum_u64(self, send, recv, root):
        recv[:] = send

    def broadcast_inplace(self, buf, root):
        pass


class GlooColl:
    def __init__(self, dist, world, rank):
        self.dist, self.world, self.rank = dist, world, rank

    def all_gather_inplace(self, buf, per):
        """buf: flat numpy array of world * per elements; this rank's slice is filled."""
        import torch
        t = torch.from_numpy(buf.view(np.uint8).reshape(self.world, -1))
        mine = t[self.rank].clone()
        self.dist.all_gather_into_tensor(t.view(-1), mine.view(-1))

    def reduce_sum_u64(self, send, recv, root):
        import torch
        t = torch.from_numpy(send.view(np.int64)).clone()      # two's-complement wrap-around == the u64 sum
        self.dist.reduce(t, dst=root, op=self.dist.ReduceOp.SUM)
        if self.rank == root:
            recv[:] = t.numpy().view(np.uint64)


    def broadcast_inplace(self, buf, root):
        import torch
        t = torch.from_numpy(buf.view(np.uint8))
        self.dist.broadcast(t, src=root)


def encrypt_sharded(blocks, coll, m, batch, nonce0):
    """blocks: tests.emu.EmuBlocks with the public key loaded.  Returns this rank's shard buffer (see nttb200_shard_plan)."""
    n, r, G, g = blocks.n, blocks.r, coll.world, coll.rank
    per = batch // G
    plan, words = shard_plan(r - 1, n, batch, G, g)
    c_shard = np.zeros(max(words, 1), dtype=np.uint64)
    ub = np.zeros(batch * n, dtype=np.uint8)
    es = np.zeros(batch * 2 * n, dtype=np.int8)
    cl = np.zeros(batch * 2 * n, dtype=np.uint64)
    own = g * per
    blocks.enc_sample(ub, None, batch, nonce0, 1, 0)
    blocks.enc_sample(None, es[own * 2 * n:], per, nonce0 + own, 0, 1)
    cl_own = cl[own * 2 * n:(own + per) * 2 * n]
    blocks.enc_front(cl_own, 1, r - 1, 1, per, ub[own * n:])
    blocks.enc_finish_last(cl_own, 2 * n, n, es[own * 2 * n:], per)
    coll.all_gather_inplace(cl, per * 2 * n)
    coll.all_gather_inplace(es, per * 2 * n)
    for (it, items, f, cnt, off) in plan:
        if cnt:
            blocks.enc_front(c_shard[off:], cnt, f, cnt, items, ub[it * n:])
    for (it, items, f, cnt, off) in plan:
        if cnt:
            blocks.enc_finish_limbs(c_shard[off:], cnt, f, cnt, items, cl[it * 2 * n:], 2 * n, n, es[it * 2 * n:], m[it * n:])
    return c_shard, es


def decrypt_sharded(blocks, coll, c_shard, batch, chunks=2):
    """blocks: EmuBlocks with the secret key loaded.  Returns m_out[batch * n] (complete on every rank).  Mode 0 of csrc/sharded.cu:
    block by block in `chunks` pieces -- partial sums -> reduce to the owner -> the owner rounds -> the owner broadcasts the block."""
    n, r, G, g = blocks.n, blocks.r, coll.world, coll.rank
    R = blocks.R
    rp, per = r - 1, batch // G
    packed = rp * (int(R.t) - 1) < 65536
    out16 = int(R.t) <= 65536
    pw = n + n // 4 if packed else 2 * n
    plan, _ = shard_plan(rp, n, batch, G, g)
    while chunks > 1 and per % chunks:
        chunks -= 1
    sub = per // chunks
    partial = np.zeros(batch * pw, dtype=np.uint64)
    recv = np.zeros(per * pw, dtype=np.uint64)
    own = g * per
    m_out = np.zeros(batch * n, dtype=np.uint64)
    plain = np.zeros(batch * n, dtype=np.uint16)
    for j, (it, items, f, cnt, off) in enumerate(plan):
        for c in range(chunks):
            pj = partial[(j * per + c * sub) * pw:(j * per + (c + 1) * sub) * pw]
            if cnt:
                blocks.dec_partial(pj, packed, c_shard[off + c * sub * 2 * cnt * n:], cnt, f, cnt, sub)
            rc = recv[c * sub * pw:(c + 1) * sub * pw]
            coll.reduce_sum_u64(pj, rc, j)
            if j == g:
                if out16:
                    blocks.dec_finish(plain[(own + c * sub) * n:], 1, rc, packed, sub)
                else:
                    blocks.dec_finish(m_out[(own + c * sub) * n:], 0, rc, packed, sub)
        if out16:
            blk = plain[it * n:(it + per) * n]
            coll.broadcast_inplace(blk, j)
            blocks.dec_expand16(blk, m_out[it * n:], per)
        else:
            coll.broadcast_inplace(m_out[it * n:(it + per) * n], j)
    return m_out


def full_from_shards(shards, n, r, batch, world):
    """Reassembles c[batch][2][r][n] (padding limb zero) from every rank's shard buffer."""
    full = np.zeros((batch, 2, r, n), dtype=np.uint64)
    for g in range(world):
        plan, _ = shard_plan(r - 1, n, batch, world, g)
        for (it, items, f, cnt, off) in plan:
            if cnt:
                full[it:it + items, :, f:f + cnt, :] = shards[g][off:off + items * 2 * cnt * n].reshape(items, 2, cnt, n)
    return full
