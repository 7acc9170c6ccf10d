"""Compute-only share of ONE rank of a world of G in the limb-sharded BFV calls (collectives skipped: nttb200_comm_fake), on one GPU:
what the transforms of a rank cost when nothing is exchanged -- the bound the 8-GPU numbers are compared with.  JSON on stdout."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))
import torch  # noqa: E402
import nttb200  # noqa: E402
from nttb200 import params  # noqa: E402

if __name__ == "__main__":
    name, total, G = "32k_16q", 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n, qs, roots = params.RNS_SETS[name]
    rn = len(qs) * n
    bfv = nttb200.Bfv(n, qs, roots)
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk)
    bfv.load_keys(sk, pk)
    m = torch.randint(0, params.T, (total * n,), dtype=torch.int64, device="cuda")
    out = {}
    for rank in (0, 3):
        comm = nttb200.Comm.fake(G, rank)
        shard = torch.zeros(bfv.shard_words(comm, total), dtype=torch.int64, device="cuda")
        res = torch.zeros(total * n, dtype=torch.int64, device="cuda")
        for mode, chunks in ((2, 4), (2, 2), (3, 1)):
            bfv.shard_config(mode, chunks)
            for _ in range(2):
                bfv.encrypt_sharded(comm, shard, m, total)
                bfv.decrypt_sharded(comm, res, shard, total)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            te = td = 0.0
            for _ in range(3):
                ev[0].record(); bfv.encrypt_sharded(comm, shard, m, total); ev[1].record(); bfv.decrypt_sharded(comm, res, shard, total); ev[2].record()
                torch.cuda.synchronize()
                te += ev[0].elapsed_time(ev[1]) / 3; td += ev[1].elapsed_time(ev[2]) / 3
            out[f"rank{rank}_mode{mode}_rounds{chunks}"] = {"encrypt_ms": round(te, 3), "decrypt_ms": round(td, 3)}
        comm.close()
        del shard, res
    print(json.dumps(out))
