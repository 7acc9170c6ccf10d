#!/usr/bin/env python
"""bench.py -- headline benchmark: batched 60-bit negacyclic NTT at N = 2^15 (BASELINE.json configs[1]).

One "step" = one forward NTT over a batch of 1024 polynomials x 32768 coefficients (256 MiB of u64, larger than the
126 MB L2, so every step streams from HBM) that cycles through the 16 demo primes (poly p uses limb p % 16).
`value` = NTT/s with the batch resident in HBM; `e2e` = the same metric through the host-buffer C-ABI entry point
(pinned host memory, H2D + D2H inside the timed region).  N > 1: one process per GPU, each rank transforms its own
batch (weak scaling, no collective on the data path), time = max over ranks.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

--impl reference times the CPU restatement of the reference's transform (oracle/, all host threads) on a bounded
sample of the same workload: the reference itself has no CPU implementation of the NTT.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))

N = 32768
POLYS = 1024
LIMBS = 16
WORKLOAD = "batched 60-bit NTT, N=2^15, 1024 polynomials, 16 RNS limbs (demo.cu 32k_16q primes)"
METRIC = "60-bit NTT/s at N=2^15 batched"
UNIT = "NTT/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------------------------
# CPU baseline (oracle port, all host threads) -- used by cpu_baseline and by --impl reference
# --------------------------------------------------------------------------------------------------------------------
def cpu_ntt_rate(budget_s=12.0, threads=None):
    import numpy as np
    from nttb200 import params
    from oracle import oracle as orc
    orc.build()
    n, qs, roots = params.RNS_SETS["32k_16q"]
    threads = threads or (os.cpu_count() or 1)
    psi = np.stack([orc.fill_psi_tables(r, q, n)[0] for q, r in zip(qs, roots)])
    qa = np.array(qs, dtype=np.uint64)
    per_thread = 16                      # one limb cycle per thread
    total = per_thread * threads
    a = np.concatenate([orc.fill_uniform(n, qs[p % LIMBS], 0x5EED0000 + p) for p in range(min(total, 64))])
    a = np.ascontiguousarray(np.resize(a, total * n))
    # one calibration transform to size the sample
    t0 = time.perf_counter()
    orc.forward_ntt_fast_range(a, n, psi, LIMBS, qa, 0, 1)
    one = time.perf_counter() - t0
    reps = max(1, int(budget_s / max(one * per_thread, 1e-6)))       # ~budget_s seconds of CPU work per thread

    def work(t):
        for _ in range(reps):
            orc.forward_ntt_fast_range(a, n, psi, LIMBS, qa, t * per_thread, (t + 1) * per_thread)

    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    done = total * reps
    return {"value": done / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{done} forward NTTs (N=2^15, same 16 primes/tables) on {threads} host threads in {dt:.1f} s; "
                      f"scalar C oracle (unsigned __int128 %), gcc -O3"}, dt, done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step_budget = max(2.0, min(20.0, 60.0 / max(args.steps + args.warmup, 1)))
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        cb, dt, done = cpu_ntt_rate(budget_s=per_step_budget)
        if i >= args.warmup:
            vals.append((done, dt))
        last = cb
    done = sum(v[0] for v in vals)
    dt = sum(v[1] for v in vals)
    value = done / dt
    last["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU arm: bounded sample per step, all host threads"},
            "cpu_baseline": last,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.proc = dev, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        # "under load" = upper half of the samples (idle samples before/after the timed region are excluded)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def bfv_throughput(nttb200, params, torch, world, dist, batch=64, reps=5):
    """BFV encrypt + decrypt ops/s (whole job): each rank encrypts and decrypts its own batch (batch sharding, no
    collective).  Decryption overwrites c1, so every repetition decrypts a fresh device copy (copy time subtracted)."""
    n, qs, roots = params.RNS_SETS["32k_16q"]
    rn = len(qs) * n
    bfv = nttb200.Bfv(n, qs, roots)
    bfv.reserve(batch)
    sk = torch.zeros(rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(2 * rn, dtype=torch.int64, device="cuda")
    bfv.keygen(sk, pk)
    m = torch.randint(0, params.T, (batch * n,), dtype=torch.int64, device="cuda")
    c = torch.zeros(batch * 2 * rn, dtype=torch.int64, device="cuda")
    keep = torch.zeros_like(c)
    out = torch.zeros(batch * n, dtype=torch.int64, device="cuda")

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    enc = timed(lambda: bfv.encrypt(c, pk, m, batch=batch))
    keep.copy_(c)

    def dec():
        c.copy_(keep)
        bfv.decrypt(out, c, sk, batch=batch)

    copy_ms = timed(lambda: c.copy_(keep))
    dec_ms = timed(dec) - copy_ms
    ok = bool(torch.equal(out, m))
    # keys loaded into the context once (nttb200_bfv_load_keys): the fused NTT (.) key -> INTT kernels
    bfv.load_keys(sk, pk)
    enc_l = timed(lambda: bfv.encrypt(c, None, m, batch=batch))
    keep.copy_(c)

    def dec_loaded():
        c.copy_(keep)
        bfv.decrypt(out, c, None, batch=batch)

    dec_l = timed(dec_loaded) - copy_ms
    ok = ok and bool(torch.equal(out, m))
    bfv.close()
    if not ok:
        raise SystemExit("bench.py: BFV round trip failed")
    return {"workload": "BFV encrypt + decrypt, n=32768, 16-limb q (demo.cu), t=1024, batch %d per GPU" % batch,
            "loaded_keys": {"enc_plus_dec_per_s": world * batch / ((enc_l + dec_l) * 1e-3), "encrypt_per_s": world * batch / (enc_l * 1e-3),
                            "decrypt_per_s": world * batch / (dec_l * 1e-3), "api": "nttb200_bfv_load_keys + encrypt / decrypt with NULL key"},
            "enc_plus_dec_per_s": world * batch / ((enc + dec_ms) * 1e-3), "encrypt_per_s": world * batch / (enc * 1e-3),
            "decrypt_per_s": world * batch / (dec_ms * 1e-3), "unit": "ops/s", "roundtrip_ok": ok}


def ncu_numbers(kernel):
    """(dram traffic per launch, fma-heavy pipe record) of `kernel` (forward, C2 workload) from the committed ncu --set full
    summary profiles/r01_final_pipe_util.json (scripts/ncu_summary.py); fallbacks are the values of the earlier capture."""
    traffic = {"ntt_strided_pass": 484.5e6, "ntt_contig_pass": 493.4e6}[kernel]
    rec = None
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r01_final_pipe_util.json")))
        for name, m in d.items():
            if kernel in name and "ShoupLazyPolicy" in name and ", 0>" in name.replace("(bool)", ""):
                traffic = m.get("traffic_bytes", traffic)
                f = m.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed")
                if f is not None:
                    rec = {"bound": "fma-heavy (IMAD/IMAD.WIDE) pipe", "frac": f / 100.0,
                           "metric": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "kernel": kernel,
                           "source": "profiles/r01_final_ncu_full_summary.md (ncu --set full of this command; not a live number)"}
    except (OSError, ValueError):
        pass
    return traffic, rec


def reference_gpu_rebuilt():
    """The reference's own kernels rebuilt for sm_100a (oracle/_ref/ref_dump), same batch, same GPU,
    in a separate process.  Reported baseline only; absent when the binary was not built."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe, "bench", "32k_16q", str(POLYS), "20"], capture_output=True, text=True, timeout=300).stdout
        d = json.loads(out.strip().splitlines()[-1])
        return {"fwd_ntt_per_s": d["fwd_ntt_per_s"], "inv_ntt_per_s": d["inv_ntt_per_s"], "keygen_us": d["keygen_us"],
                "encrypt_us": d["encrypt_us"], "decrypt_us": d["decrypt_us"],
                "note": "forwardNTT_batch / inverseNTT_batch num=1024 and single-item keygen/encryption/decryption_rns loops, CUDA events"}
    except Exception as e:  # pragma: no cover
        return {"error": str(e)[:200]}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import nttb200
    from nttb200 import params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, qs, roots = params.RNS_SETS["32k_16q"]
    ctx = nttb200.Context(n, qs, roots)
    # synthetic batch: uniform residues, poly p in [0, q_{p % 16})
    g = torch.Generator(device="cuda").manual_seed(0x5EED0000 + rank)
    qv = torch.tensor(qs, dtype=torch.int64, device="cuda").repeat(POLYS // LIMBS).view(POLYS, 1)
    a = torch.randint(0, 2**62, (POLYS, n), dtype=torch.int64, device="cuda", generator=g) % qv
    a0 = a.clone()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness gate before timing: INTT(NTT(a)) == a on the device (oracle parity is tests/'s job)
    ctx.forward_ntt_batch(a, POLYS, LIMBS)
    ctx.inverse_ntt_batch(a, POLYS, LIMBS)
    if not torch.equal(a, a0):
        raise SystemExit("bench.py: round-trip check failed -- refusing to time a wrong kernel")

    for _ in range(args.warmup):
        ctx.forward_ntt_batch(a, POLYS, LIMBS)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    # ---- timed region: K steps; each step = pass 1 + pass 2, with an event between them so the dominant kernel's
    # duration is measured live (events are ~us, kernels ~100s of us)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        ev[k][0].record()
        ctx.ntt_pass(a, POLYS, LIMBS, False, 0)
        ev[k][1].record()
        ctx.ntt_pass(a, POLYS, LIMBS, False, 1)
        ev[k][2].record()
    t_end.record()
    barrier()
    # the timed region is only a few milliseconds long (nvidia-smi samples every 100 ms): keep the SAME launches running for another
    # 0.4 s, untimed, so that the clock / throttle record really is one taken under this load
    t_hold = time.perf_counter()
    while time.perf_counter() - t_hold < 0.4:
        for _ in range(50):
            ctx.ntt_pass(a, POLYS, LIMBS, False, 0)
            ctx.ntt_pass(a, POLYS, LIMBS, False, 1)
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed region + 0.4 s untimed continuation of the same launches"
    ms_total = t_start.elapsed_time(t_end)
    p1 = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    p2 = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * POLYS * args.steps / (ms_total * 1e-3)

    # ---- inverse transform, same protocol (reported as an extra)
    for _ in range(2):
        ctx.inverse_ntt_batch(a, POLYS, LIMBS)
    barrier()
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record()
    for _ in range(args.steps):
        ctx.inverse_ntt_batch(a, POLYS, LIMBS)
    i1.record()
    barrier()
    inv_ms = i0.elapsed_time(i1) / args.steps

    # ---- end to end through the host-buffer C-ABI entry point (pinned host memory)
    hin = torch.empty((POLYS, n), dtype=torch.int64, pin_memory=True)
    hout = torch.empty((POLYS, n), dtype=torch.int64, pin_memory=True)
    hin.copy_(a0.cpu())
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        ctx.forward_ntt_batch_host(hin.numpy(), hout.numpy(), POLYS, LIMBS)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.forward_ntt_batch_host(hin.numpy(), hout.numpy(), POLYS, LIMBS)   # synchronous
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_value = world * POLYS * e2e_steps / e2e_s
    # spot check of the e2e result against the device path
    chk = a0[:2].clone()
    ctx.forward_ntt_batch(chk, 2, LIMBS)
    if not torch.equal(chk.cpu(), hout[:2]):
        raise SystemExit("bench.py: host-buffer path disagrees with the device path")

    # ---- secondary metric of BASELINE.json ("BFV enc+dec ops/s"): batched context API, demo.cu's (32768, 16 primes) set
    bfv_line = None
    if not args.no_bfv:
        bfv_line = bfv_throughput(nttb200, params, torch, world, dist)

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = 16.0 * n * POLYS                     # per launch: every coefficient read once + written once
        dom, dom_ms = ("ntt_strided_pass", p1) if p1 >= p2 else ("ntt_contig_pass", p2)
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this workload
        # (profiles/r01z_ncu_full_summary.md); below the algorithmic 536.9 MB because write-back still sits in L2 at kernel end
        ncu_traffic, int_pipe = ncu_numbers(dom)
        ach = alg_bytes / (dom_ms * 1e-3) / 1e9
        butterflies = POLYS * (n // 2) * 15
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _ = cpu_ntt_rate(budget_s=12.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "polys_per_gpu": POLYS, "n": n, "limbs": LIMBS,
                       "l2": "batch is 256 MiB per GPU (> 126 MB L2): every step streams from HBM, no flush needed",
                       "parallelism": f"batch-sharded x{world}, no collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": POLYS * n * 8, "d2h_bytes_per_step": POLYS * n * 8,
                    "steps": e2e_steps, "api": "nttb200_forward_ntt_batch_host (pinned host buffers, 3-stage stream pipeline)"},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": ncu_traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": dom_ms},
            # the pipe that actually binds these kernels: ncu's own utilisation counter of the fma-heavy
            # (IMAD / IMAD.WIDE) pipe for the same kernel and workload, from the committed --set full capture (not measured live)
            "roofline_int_pipe": int_pipe,
            "kernels_ms": {"ntt_strided_pass": p1, "ntt_contig_pass": p2},
            "hbm_gbs_whole_step": 2 * alg_bytes / ((p1 + p2) * 1e-3) / 1e9,
            "butterflies_per_s": butterflies * args.steps / (ms_total * 1e-3) * world,
            "inverse": {"value": world * POLYS / (inv_ms * 1e-3), "unit": "INTT/s", "ms_per_step": inv_ms},
            "cpu_baseline": cb,
            "bfv": bfv_line,
            "reference_gpu_rebuilt": reference_gpu_rebuilt() if world == 1 and not args.no_cpu_baseline else None,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bfv", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
