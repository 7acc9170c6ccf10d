"""Loaded-key encryption (fused epilogue on / off) and decryption, once each after warm-up, for an ncu launch list:
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file out.csv \
       python scripts/enc_launches.py [--batch 64] [--set 32k_16q]
The profiled region is bracketed by cudaProfilerStart/Stop (run ncu with --profile-from-start off)."""
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
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    c = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    m = torch.randint(0, params.T, (B * n,), dtype=torch.int64, device="cuda")
    out = torch.zeros(B * n, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk)
    bfv.load_keys(sk, pk)
    for fused in (True, False):
        bfv.set_fused_epilogue(fused)
        bfv.encrypt(c, None, m, batch=B)
    bfv.decrypt(out, c, None, batch=B)
    assert torch.equal(out, m)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    bfv.set_fused_epilogue(True)
    bfv.encrypt(c, None, m, batch=B)
    torch.cuda.synchronize()
    bfv.set_fused_epilogue(False)
    bfv.encrypt(c, None, m, batch=B)
    torch.cuda.synchronize()
    bfv.decrypt(out, c, None, batch=B)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    assert torch.equal(out, m)
    print("ok")
