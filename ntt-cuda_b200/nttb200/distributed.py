"""Multi-GPU orchestration: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing.

Round 2: the limb-sharded BFV calls live in the C library (nttb200_bfv_encrypt_sharded / _decrypt_sharded, csrc/sharded.cu: balanced
(limb, block) tiles, peer-to-peer exchange over CUDA IPC with NCCL fallback); from Python they are `Bfv.encrypt_sharded` /
`Bfv.decrypt_sharded` with a `nttb200.Comm` (Comm.from_torch() adopts the process group's NCCL communicator).  The helpers below are
the round-1 building blocks (contiguous limb ranges, caller-side all-reduce) kept for callers that run their own collectives.

The hot path shards with NO data-path collective for NTT / INTT / pointwise / keygen / encryption (units = (batch item,
limb) are independent) and with exactly ONE collective for decryption: the cross-limb base-conversion sum.

  * shard_batch(total, world, rank)             -> the contiguous slice of batch items a rank owns (batch sharding)
  * shard_limbs(rp, world, rank)                -> the contiguous limb range a rank owns (limb sharding)
  * scatter_ciphertext_limbs / decrypt_limb_sharded : limb-sharded decryption, all-reduce(SUM) of [batch][2][n] u64 partials

`backend` hooks let the CPU test-suite (gloo, world_size 2) drive the identical control flow with the emulator build of
the kernels in place of the GPU library.
"""
from __future__ import annotations


def shard_batch(total: int, world: int, rank: int):
    """Contiguous, balanced split: the first (total % world) ranks get one extra item.  Returns (first, count)."""
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def shard_limbs(rp: int, world: int, rank: int):
    """Limb-major blocks of ceil(rp / world) limbs.  Ranks beyond the last limb own nothing."""
    per = -(-rp // world)
    first = min(rank * per, rp)
    return first, max(0, min(per, rp - first))


def ciphertext_limb_shard(c, n: int, r: int, first: int, count: int, batch: int = 1):
    """Slices a reference-layout ciphertext array c[batch][2][r][n] (numpy or torch, flat or shaped) into the compact
    shard [batch][2][count][n] of limbs [first, first+count)."""
    v = c.reshape(batch, 2, r, n)
    s = v[:, :, first:first + count, :]
    return s.contiguous().reshape(-1) if hasattr(s, "contiguous") else s.copy().reshape(-1)


def sub_ring(q, psi_roots, first: int, count: int):
    """Moduli / roots of the sub-ring a rank encrypts on: its owned limbs followed by the last limb (which the modulus
    switch drops).  A ciphertext limb depends only on itself, on the last limb and on randomness whose keystream layout
    (9n bytes, bfv_encryption.cuh:228) is independent of the limb count."""
    q, psi_roots = list(q), list(psi_roots)
    return q[first:first + count] + [q[-1]], psi_roots[first:first + count] + [psi_roots[-1]]


def public_key_limb_shard(pk, n: int, r: int, first: int, count: int):
    """pk[2][r][n] -> [2][count+1][n]: owned limbs then the last limb, per half."""
    v = pk.reshape(2, r, n)
    idx = list(range(first, first + count)) + [r - 1]
    s = v[:, idx, :]
    return s.contiguous().reshape(-1) if hasattr(s, "contiguous") else s.copy().reshape(-1)


def encrypt_limb_sharded(sub_bfv, c_shard, pk_shard, m, batch: int, nonce0: int = 0):
    """Limb-sharded encryption on this rank, no communication: `sub_bfv` is a Bfv on sub_ring(...); c_shard[batch][2][count+1][n]
    receives the owned limbs of every ciphertext (slot `count` of each half is the padding limb)."""
    sub_bfv.encrypt(c_shard, pk_shard, m, batch=batch, nonce0=nonce0)
    return c_shard


def decrypt_limb_sharded(bfv, c_shard, sk_shard, first: int, count: int, batch: int, all_reduce_sum, new_u64, sk_per_item=False,
                         shard_half_limbs: int = 0):
    """Limb-sharded decryption on this rank.

    bfv            object with decrypt_partial / decrypt_finish (nttb200.Bfv on the GPU)
    all_reduce_sum callable(buffer) -> None, in-place 64-bit SUM over all ranks (torch.distributed.all_reduce on the
                   int64 view: two's-complement wrap-around is exactly the u64 sum the kernels need)
    new_u64        callable(count) -> zero-initialised device buffer of `count` 64-bit words
    Returns the plaintext buffer m_out[batch][n] (identical on every rank)."""
    n = bfv.n
    gamma, world = getattr(bfv, "gamma", None), getattr(bfv, "world", None)
    if gamma and world:          # the 64-bit SUM of `world` partial sums below gamma must not wrap
        assert world * gamma < 1 << 64, "all_reduce(SUM) of the gamma partials would wrap: reduce in two levels"
    partial = new_u64(batch * 2 * n)
    if count > 0:
        bfv.decrypt_partial(partial, c_shard, sk_shard, first, count, batch=batch, sk_per_item=sk_per_item, shard_half_limbs=shard_half_limbs)
    all_reduce_sum(partial)                      # the path's only collective: batch * 2 * n * 8 bytes
    out = new_u64(batch * n)
    bfv.decrypt_finish(out, partial, batch=batch)
    return out


def torch_all_reduce_sum(t):
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def torch_new_u64(count, device="cuda"):
    import torch
    return torch.zeros(count, dtype=torch.int64, device=device)
