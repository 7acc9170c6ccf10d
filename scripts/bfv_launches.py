"""One batched keygen / encrypt / decrypt (+ loaded-key variants) for an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/bfv_launches.py [--batch 64]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))

if __name__ == "__main__":
    import torch
    import nttb200
    from nttb200 import params
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="32k_16q")
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    n, qs, roots = params.RNS_SETS[a.set]
    rn = len(qs) * n
    B = a.batch
    bfv = nttb200.Bfv(n, qs, roots)
    bfv.reserve(B)
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    skb = torch.zeros(B * rn, dtype=torch.int64, device="cuda")
    pkb = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    c = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    m = torch.randint(0, params.T, (B * n,), dtype=torch.int64, device="cuda")
    out = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    bfv.keygen(skb, pkb, batch=B)
    sk.copy_(skb[:rn]); pk.copy_(pkb[:2 * rn])
    bfv.encrypt(c, pk, m, batch=B)
    bfv.decrypt(out, c, sk, batch=B)
    assert torch.equal(out, m)
    bfv.load_keys(sk, pk)
    bfv.encrypt(c, None, m, batch=B)
    bfv.decrypt(out, c, None, batch=B)
    assert torch.equal(out, m)
    torch.cuda.synchronize()
    print("ok")
