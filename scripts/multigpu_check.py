"""Run under torchrun on N GPUs (one process per GPU): the limb-sharded BFV calls of the C library (NCCL inside libnttb200.so) must be
bit-identical to the single-GPU calls, for both collective modes; prints one JSON line on rank 0 with timings.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py
  env: NTTB200_SET (default 32k_16q), NTTB200_BATCH (default 8 * world), NTTB200_OWN_COMM=1 (communicator from a broadcast unique id
       instead of torch's), NTTB200_REPS
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import nttb200  # noqa: E402
from nttb200 import params  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = os.environ.get("NTTB200_SET", "32k_16q")
    n, qs, roots = params.RNS_SETS[name]
    r = len(qs)
    rn = r * n
    B = int(os.environ.get("NTTB200_BATCH", str(8 * world)))
    reps = int(os.environ.get("NTTB200_REPS", "5"))
    bfv = nttb200.Bfv(n, qs, roots)
    if os.environ.get("NTTB200_OWN_COMM") == "1":
        def bcast(raw):
            t = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if raw is not None:
                t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
            dist.broadcast(t, 0)
            return bytes(t.cpu().numpy())
        comm = nttb200.Comm.create(world, rank, bcast)
    else:
        comm = nttb200.Comm.from_torch()
    # every rank derives the same key pair deterministically (nonce-addressed sampling: no broadcast needed)
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk)
    bfv.load_keys(sk, pk)
    g = torch.Generator(device="cuda").manual_seed(7)
    m = torch.randint(0, params.T, (B * n,), dtype=torch.int64, device="cuda", generator=g)
    # single-GPU results (every rank computes them: the comparison needs no communication)
    c_full = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    bfv.encrypt(c_full, None, m, batch=B, nonce0=3)
    ref = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    bfv.decrypt(ref, c_full.clone(), None, batch=B)
    ok_single = bool(torch.equal(ref, m))
    words = bfv.shard_words(comm, B)
    expect_shard = torch.zeros(max(words, 1), dtype=torch.int64, device="cuda")
    bfv.shard_from_full(world, rank, expect_shard, c_full, B)
    shard = torch.zeros_like(expect_shard)
    res = {}
    ok_enc = ok_dec = True
    for mode, chunks in ((4, 0), (4, 2), (2, 4), (3, 1), (0, 4), (1, 4)):      # (4, 0) = the library's default schedule
        bfv.shard_config(mode, chunks)
        shard.zero_()
        bfv.encrypt_sharded(comm, shard, m, B, nonce0=3)
        ok_enc = ok_enc and bool(torch.equal(shard, expect_shard))
        out = torch.zeros(B * n, dtype=torch.int64, device="cuda")
        bfv.decrypt_sharded(comm, out, shard, B)
        ok_dec = ok_dec and bool(torch.equal(out, ref))
        # timing: encrypt + decrypt pairs (decrypt consumes what encrypt wrote)
        for _ in range(2):
            bfv.encrypt_sharded(comm, shard, m, B, nonce0=3)
            bfv.decrypt_sharded(comm, out, shard, B)
        torch.cuda.synchronize()
        dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        te = td = 0.0
        for _ in range(reps):
            ev[0].record()
            bfv.encrypt_sharded(comm, shard, m, B, nonce0=3)
            ev[1].record()
            bfv.decrypt_sharded(comm, out, shard, B)
            ev[2].record()
            torch.cuda.synchronize()
            te += ev[0].elapsed_time(ev[1])
            td += ev[1].elapsed_time(ev[2])
        t = torch.tensor([te / reps, td / reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok_dec = ok_dec and bool(torch.equal(out, ref))
        res[f"mode{mode}_chunks{chunks}"] = {"encrypt_ms": float(t[0]), "decrypt_ms": float(t[1]),
                                            "enc_plus_dec_per_s": B / ((float(t[0]) + float(t[1])) * 1e-3)}
    # single-GPU time of the same batch on this box (rank 0's GPU; all ranks run it, max taken)
    for _ in range(2):
        bfv.encrypt(c_full, None, m, batch=B, nonce0=3)
        bfv.decrypt(ref, c_full, None, batch=B)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        bfv.encrypt(c_full, None, m, batch=B, nonce0=3)
        bfv.decrypt(ref, c_full, None, batch=B)
    e1.record()
    torch.cuda.synchronize()
    t1 = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t1, op=dist.ReduceOp.MAX)
    flags = torch.tensor([int(ok_single), int(ok_enc), int(ok_dec)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        single = B / (float(t1[0]) * 1e-3)
        for v in res.values():
            v["speedup_vs_one_gpu"] = v["enc_plus_dec_per_s"] / single
        print(json.dumps({"world": world, "set": name, "batch": B, "single_gpu_round_trip_ok": bool(flags[0].item()),
                          "sharded_encrypt_bit_identical": bool(flags[1].item()), "sharded_decrypt_bit_identical": bool(flags[2].item()),
                          "one_gpu_enc_plus_dec_per_s": single, "sharded": res,
                          "tiles_per_rank": [sum(b[3] for b in nttb200.shard_plan(r - 1, n, B, world, k)[0]) for k in range(world)],
                          "comm": "own (unique id)" if os.environ.get("NTTB200_OWN_COMM") == "1" else "adopted from torch.distributed"}))
    comm.close()
    bfv.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
