"""A/B of library builds on the SAME box: forward / inverse NTT at C2 (1024 x 2^15, 16 primes), each build in its own process, interleaved.
   python scripts/ab_ntt.py libA.so libB.so [...]   (paths relative to ntt-cuda_b200/nttb200/)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "ntt-cuda_b200"))
import torch, nttb200
from nttb200 import params
n, qs, roots = params.RNS_SETS["32k_16q"]
ctx = nttb200.Context(n, qs, roots)
P = 1024
qv = torch.tensor(qs, dtype=torch.int64, device="cuda").repeat(P // 16).view(P, 1)
a = torch.randint(0, 2**62, (P, n), dtype=torch.int64, device="cuda") %% qv
out = {}
b = a.clone()
ctx.forward_ntt_batch(b, P, 16); ctx.inverse_ntt_batch(b, P, 16)
out["roundtrip_ok"] = bool(torch.equal(a, b))
for name, fn in (("fwd_strided", lambda: ctx.ntt_pass(a, P, 16, False, 0)), ("fwd_contig", lambda: ctx.ntt_pass(a, P, 16, False, 1)),
                 ("inv_contig", lambda: ctx.ntt_pass(a, P, 16, True, 0)), ("inv_strided", lambda: ctx.ntt_pass(a, P, 16, True, 1))):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): fn()
    e1.record(); torch.cuda.synchronize()
    out[name] = round(e0.elapsed_time(e1) / 200 * 1000, 2)
print(json.dumps(out))
''' % (ROOT, ROOT)

if __name__ == "__main__":
    libs = sys.argv[1:]
    for rep in range(3):
        for lib in libs:
            env = dict(os.environ, NTTB200_LIB=os.path.join(ROOT, "ntt-cuda_b200", "nttb200", lib))
            o = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
            print(lib, rep, o.stdout.strip() or o.stderr[-300:])
