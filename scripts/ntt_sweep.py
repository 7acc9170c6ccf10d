"""Forward / inverse NTT throughput for every supported ring degree (batch sized to 256 MiB, 8 limbs or as available)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))
import torch, nttb200
from nttb200 import params
out = []
for logn in range(11, 18):
    n = 1 << logn
    limbs = 8
    qs, roots = params.find_ntt_primes(55, n, limbs)
    ctx = nttb200.Context(n, qs, roots)
    num = (256 << 20) // (n * 8)
    qv = torch.tensor(qs, dtype=torch.int64, device="cuda").repeat(num // limbs).view(num, 1)
    a = torch.randint(0, 2**62, (num, n), dtype=torch.int64, device="cuda") % qv
    res = {"logn": logn, "num": num}
    for inv in (False, True):
        fn = (lambda: ctx.inverse_ntt_batch(a, num, limbs)) if inv else (lambda: ctx.forward_ntt_batch(a, num, limbs))
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        key = "inv" if inv else "fwd"
        res[key + "_ms"] = round(ms, 4)
        res[key + "_per_s"] = round(num / (ms * 1e-3))
        res[key + "_Gbfly_s"] = round(num * (n // 2) * logn / (ms * 1e-3) / 1e9, 1)
        res[key + "_hbm_GBs_2pass"] = round(2 * 16 * n * num / (ms * 1e-3) / 1e9)
    out.append(res); ctx.close()
    print(json.dumps(res))
