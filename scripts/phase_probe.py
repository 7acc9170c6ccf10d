"""Profiling aid: times each NTT kernel normally, with the butterflies skipped and with the tile traffic skipped
(NTTB200_DEBUG_SKIP is read when the context is created).  Results of the skipping modes are of course wrong."""
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
POLYS, L = 1024, 16
qv = torch.tensor(qs, dtype=torch.int64, device="cuda").repeat(POLYS // L).view(POLYS, 1)
a = torch.randint(0, 2**62, (POLYS, n), dtype=torch.int64, device="cuda") % qv
out = {}
for mode, name in ((0, "normal"), (1, "no_compute"), (2, "no_memory")):
    os.environ["NTTB200_DEBUG_SKIP"] = str(mode)
    ctx = nttb200.Context(n, qs, roots)
    res = {}
    for inv in (False, True):
        for which in (0, 1):
            for _ in range(3):
                ctx.ntt_pass(a, POLYS, L, inv, which)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ctx.ntt_pass(a, POLYS, L, inv, which)
            e1.record()
            torch.cuda.synchronize()
            res[("inv" if inv else "fwd") + "_pass%d" % which] = round(e0.elapsed_time(e1) / 20, 4)
    out[name] = res
    ctx.close()
    a %= qv
print(json.dumps(out))
