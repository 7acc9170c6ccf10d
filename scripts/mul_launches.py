"""One ciphertext x ciphertext multiplication + relinearisation (batch 8, 32768 x 16 limbs) after warm-up, for an ncu launch list."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))

if __name__ == "__main__":
    import time
    import torch
    import nttb200
    from nttb200 import params
    n, qs, roots = params.RNS_SETS["32k_16q"]
    rn = len(qs) * n
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    bfv = nttb200.Bfv(n, qs, roots)
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk)
    bfv.load_keys(sk, pk)
    bfv.relin_keygen(sk)
    m = torch.randint(0, params.T, (B * n,), dtype=torch.int64, device="cuda")
    ca = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
    cb = torch.zeros_like(ca)
    bfv.encrypt(ca, None, m, batch=B, nonce0=1)
    bfv.encrypt(cb, None, m, batch=B, nonce0=1000)
    out = torch.zeros_like(ca)
    bfv.mul(out, ca, cb, batch=B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        bfv.mul(out, ca, cb, batch=B)
    e1.record()
    torch.cuda.synchronize()
    print("mul+relin: %.3f ms per batch of %d -> %.1f products/s" % (e0.elapsed_time(e1) / 3, B, B / (e0.elapsed_time(e1) / 3 * 1e-3)))
    torch.cuda.cudart().cudaProfilerStart()
    bfv.mul(out, ca, cb, batch=B)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
