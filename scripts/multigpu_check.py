"""Run under torchrun on N GPUs: limb-sharded BFV decryption (ONE NCCL all-reduce) must equal single-GPU decryption and
the messages; batch-sharded NTT must equal the unsharded transform.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import nttb200  # noqa: E402
from nttb200 import params  # noqa: E402
from nttb200.distributed import (ciphertext_limb_shard, decrypt_limb_sharded, encrypt_limb_sharded, public_key_limb_shard,  # noqa: E402
                                 shard_batch, shard_limbs, sub_ring, torch_all_reduce_sum, torch_new_u64)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = os.environ.get("NTTB200_SET", "32k_16q")
    n, qs, roots = params.RNS_SETS[name]
    r, rp = len(qs), len(qs) - 1
    rn = r * n
    B = int(os.environ.get("NTTB200_BATCH", "32"))
    bfv = nttb200.Bfv(n, qs, roots)
    # every rank derives the same key pair / ciphertexts deterministically (nonce-addressed sampling: no broadcast needed)
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk)
    g = torch.Generator(device="cuda").manual_seed(7)
    m = torch.randint(0, params.T, (B * n,), dtype=torch.int64, device="cuda", generator=g)
    c = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    bfv.encrypt(c, pk, m, batch=B)
    # single-GPU reference result
    ref = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    bfv.decrypt(ref, c.clone(), sk, batch=B)
    # limb-sharded
    first, count = shard_limbs(rp, world, rank)
    c_shard = ciphertext_limb_shard(c, n, r, first, count, batch=B)
    sk_shard = sk.view(r, n)[first:first + count].contiguous().view(-1)
    torch.cuda.synchronize(); dist.barrier()
    reps = 5
    shard_keep = c_shard.clone()
    for _ in range(2):
        out = decrypt_limb_sharded(bfv, c_shard.copy_(shard_keep), sk_shard, first, count, B, torch_all_reduce_sum, torch_new_u64)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = decrypt_limb_sharded(bfv, c_shard.copy_(shard_keep), sk_shard, first, count, B, torch_all_reduce_sum, torch_new_u64)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok_limb = bool(torch.equal(out, ref)) and bool(torch.equal(out, m))
    # limb-sharded ENCRYPTION on the sub-ring {owned limbs, last limb}: no communication; must equal the same limbs of c,
    # and its (count+1)-limb shard must feed the limb-sharded decryption directly
    ok_enc = True
    if count > 0:
        q_sub, roots_sub = sub_ring(qs, roots, first, count)
        sub = nttb200.Bfv(n, q_sub, roots_sub)
        c_enc = torch.zeros(B * 2 * (count + 1) * n, dtype=torch.int64, device="cuda")
        encrypt_limb_sharded(sub, c_enc, public_key_limb_shard(pk, n, r, first, count), m, B)
        ok_enc = bool(torch.equal(c_enc.view(B, 2, count + 1, n)[:, :, :count, :], c.view(B, 2, r, n)[:, :, first:first + count, :]))
        out2 = decrypt_limb_sharded(bfv, c_enc, sk_shard, first, count, B, torch_all_reduce_sum, torch_new_u64, shard_half_limbs=count + 1)
        sub.close()
    else:
        out2 = decrypt_limb_sharded(bfv, None, None, first, count, B, torch_all_reduce_sum, torch_new_u64)
    ok_enc = ok_enc and bool(torch.equal(out2, m))
    # batch-sharded decryption: no collective at all
    f, cnt = shard_batch(B, world, rank)
    outb = torch.zeros(max(cnt, 1) * n, dtype=torch.int64, device="cuda")
    if cnt:
        bfv.decrypt(outb, c[f * 2 * rn:(f + cnt) * 2 * rn].clone(), sk, batch=cnt)
    ok_batch = cnt == 0 or bool(torch.equal(outb[:cnt * n], m[f * n:(f + cnt) * n]))
    flags = torch.tensor([int(ok_limb), int(ok_batch), int(ok_enc)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "set": name, "batch": B, "limb_sharded_decrypt_ok": bool(flags[0].item()),
                          "batch_sharded_decrypt_ok": bool(flags[1].item()), "limb_sharded_encrypt_then_decrypt_ok": bool(flags[2].item()),
                          "limb_sharded_decrypt_ms": float(ms.item()),
                          "limb_sharded_decrypt_per_s": B / (float(ms.item()) * 1e-3), "allreduce_bytes": B * 2 * n * 8,
                          "limbs_per_rank": [shard_limbs(rp, world, k)[1] for k in range(world)]}))
    bfv.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
