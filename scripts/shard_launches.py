"""One rank's share (rank 3 of 8, fake communicator: collectives skipped) of the limb-sharded encrypt + decrypt of 4096 ciphertexts, once,
for an ncu launch list (run ncu with --profile-from-start off)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))
import torch  # noqa: E402
import nttb200  # noqa: E402
from nttb200 import params  # noqa: E402

if __name__ == "__main__":
    n, qs, roots = params.RNS_SETS["32k_16q"]
    rn = len(qs) * n
    total, G, rank = 4096, 8, 3
    bfv = nttb200.Bfv(n, qs, roots)
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk)
    bfv.load_keys(sk, pk)
    m = torch.randint(0, params.T, (total * n,), dtype=torch.int64, device="cuda")
    comm = nttb200.Comm.fake(G, rank)
    shard = torch.zeros(bfv.shard_words(comm, total), dtype=torch.int64, device="cuda")
    res = torch.zeros(total * n, dtype=torch.int64, device="cuda")
    bfv.shard_config(int(sys.argv[1]) if len(sys.argv) > 1 else 3, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    bfv.encrypt_sharded(comm, shard, m, total)
    bfv.decrypt_sharded(comm, res, shard, total)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    bfv.encrypt_sharded(comm, shard, m, total)
    torch.cuda.synchronize()
    bfv.decrypt_sharded(comm, res, shard, total)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("ok")
