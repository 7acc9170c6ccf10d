"""Host-buffer (end-to-end) throughput of nttb200_forward_ntt_batch_host vs chunk size (NTTB200_E2E_CHUNK_MB)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))
import torch, nttb200
from nttb200 import params
n, qs, roots = params.RNS_SETS["32k_16q"]
ctx = nttb200.Context(n, qs, roots)
P = 1024
hin = torch.randint(0, 2**50, (P, n), dtype=torch.int64).pin_memory()
hout = torch.empty((P, n), dtype=torch.int64, pin_memory=True)
for _ in range(2): ctx.forward_ntt_batch_host(hin.numpy(), hout.numpy(), P, 16)
t0 = time.perf_counter()
for _ in range(8): ctx.forward_ntt_batch_host(hin.numpy(), hout.numpy(), P, 16)
dt = (time.perf_counter() - t0) / 8
# plain copies for reference
d = torch.empty((P, n), dtype=torch.int64, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(8): d.copy_(hin, non_blocking=True)
torch.cuda.synchronize(); h2d = (time.perf_counter() - t0) / 8
t0 = time.perf_counter()
for _ in range(8): hout.copy_(d, non_blocking=True)
torch.cuda.synchronize(); d2h = (time.perf_counter() - t0) / 8
print(json.dumps({"chunk_mb": os.environ.get("NTTB200_E2E_CHUNK_MB", "16"), "ntt_per_s": P / dt, "ms": dt * 1e3, "h2d_GBs": P * n * 8 / h2d / 1e9, "d2h_GBs": P * n * 8 / d2h / 1e9}))
# both directions at once (two streams): the ceiling of any H2D / compute / D2H pipeline on this host
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty((P, n), dtype=torch.int64, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(8):
    with torch.cuda.stream(s1): d.copy_(hin, non_blocking=True)
    with torch.cuda.stream(s2): hout.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); bi = (time.perf_counter() - t0) / 8
print(json.dumps({"bidirectional_GBs_each_way": P * n * 8 / bi / 1e9, "ntt_per_s_ceiling": P / bi}))
