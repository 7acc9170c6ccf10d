"""SHA-256 of nttb200_bfv_mul_tensor / _mul outputs for fixed inputs (deterministic keys and sampling): an A/B anchor for kernel
changes -- the digests must not move.  Usage: python scripts/mul_hash.py [set ...]"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))

if __name__ == "__main__":
    import torch
    import nttb200
    from nttb200 import params
    for name in (sys.argv[1:] or ["4k_3q", "16k_9q", "32k_16q"]):
        n, qs, roots = params.RNS_SETS[name]
        r = len(qs)
        rn = r * n
        B = 3
        bfv = nttb200.Bfv(n, qs, roots)
        sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
        pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
        bfv.keygen(sk, pk)
        bfv.load_keys(sk, pk)
        bfv.relin_keygen(sk)
        g = torch.Generator(device="cuda").manual_seed(5)
        m = torch.randint(0, params.T, (2 * B * n,), dtype=torch.int64, device="cuda", generator=g)
        ca = torch.zeros(B * 2 * rn, dtype=torch.int64, device="cuda")
        cb = torch.zeros_like(ca)
        bfv.encrypt(ca, None, m[:B * n], batch=B, nonce0=1)
        bfv.encrypt(cb, None, m[B * n:], batch=B, nonce0=1000)
        y = torch.zeros(B * 3 * (r - 1) * n, dtype=torch.int64, device="cuda")
        bfv.mul_tensor(y, ca, cb, batch=B)
        out = torch.zeros_like(ca)
        bfv.mul(out, ca, cb, batch=B)
        sq = torch.zeros_like(ca)
        bfv.mul(sq, ca, ca, batch=B)
        torch.cuda.synchronize()
        h = [hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()[:16] for t in (y, out, sq)]
        print(name, "tensor", h[0], "mul", h[1], "square", h[2])
        bfv.close()
