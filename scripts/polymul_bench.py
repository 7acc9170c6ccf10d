"""Fused polynomial product (nttb200_poly_mul_batch) against the same product done with separate calls
(forward_ntt_batch x2, barrett_batch, inverse_ntt_batch) at the C2 size.  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))
import torch  # noqa: E402

import nttb200  # noqa: E402
from nttb200 import params  # noqa: E402

n, qs, roots = params.RNS_SETS["32k_16q"]
L, num = 16, 1024
ctx = nttb200.Context(n, qs, roots)
qv = torch.tensor(qs, dtype=torch.int64, device="cuda").repeat(num // L).view(num, 1)
a0 = torch.randint(0, 2**62, (num, n), dtype=torch.int64, device="cuda") % qv
b0 = torch.randint(0, 2**62, (num, n), dtype=torch.int64, device="cuda") % qv
a, b = a0.clone(), b0.clone()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def unfused():
    ctx.forward_ntt_batch(a, num, L)
    ctx.forward_ntt_batch(b, num, L)
    nttb200.barrett_batch(a, b, n, num, L, ctx.q_dev, ctx.mu_dev, ctx.qbit_dev)
    ctx.inverse_ntt_batch(a, num, L)


def fused():
    ctx.poly_mul_batch(a, b, num, L)


# values drift over repetitions (products of products) but stay canonical residues: timing only
t_unf = timed(unfused)
a.copy_(a0); b.copy_(b0)
t_f = timed(fused)
print(json.dumps({"workload": "1024 polynomial products, N=2^15, 16 limbs", "unfused_ms": t_unf, "fused_ms": t_f,
                  "unfused_products_per_s": num / t_unf * 1e3, "fused_products_per_s": num / t_f * 1e3, "speedup": t_unf / t_f,
                  "launches": {"unfused": 7, "fused": 4}}))
