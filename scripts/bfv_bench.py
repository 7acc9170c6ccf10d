"""BFV keygen / encrypt / decrypt throughput of the batched context API (synthetic messages).  Prints one JSON line.
   python scripts/bfv_bench.py [--set 32k_16q] [--batch 128] [--iters 10]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))


def run(setname, batch, iters):
    import torch
    import nttb200
    from nttb200 import params
    n, qs, roots = params.RNS_SETS[setname]
    r = len(qs)
    rn = r * n
    bfv = nttb200.Bfv(n, qs, roots)
    bfv.reserve(batch)
    sk = torch.zeros(batch * rn, dtype=torch.int64, device="cuda")
    pk = torch.zeros(batch * 2 * rn, dtype=torch.int64, device="cuda")
    c = torch.zeros(batch * 2 * rn, dtype=torch.int64, device="cuda")
    c_keep = torch.zeros_like(c)
    m = torch.randint(0, params.T, (batch * n,), dtype=torch.int64, device="cuda")
    out = torch.zeros(batch * n, dtype=torch.int64, device="cuda")

    def timed(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    kg = timed(lambda: bfv.keygen(sk, pk, batch=batch), iters)
    enc = timed(lambda: bfv.encrypt(c, pk[:2 * rn], m, batch=batch), iters)
    c_keep.copy_(c)

    def dec():
        c.copy_(c_keep)          # decryption overwrites c1; the restore is part of the timed loop (device copy, ~1/10 of the work)
        bfv.decrypt(out, c, sk[:rn], batch=batch)

    dec_ms = timed(dec, iters)
    cp_ms = timed(lambda: c.copy_(c_keep), iters)
    ok = bool(torch.equal(out, m))
    # loaded-key path: fused "NTT (.) key -> INTT" kernel
    bfv.load_keys(sk[:rn], pk[:2 * rn])
    enc_f = timed(lambda: bfv.encrypt(c, None, m, batch=batch), iters)
    ok = ok and bool(torch.equal(c, c_keep))

    def dec_fused():
        c.copy_(c_keep)
        bfv.decrypt(out, c, None, batch=batch)

    out.zero_()
    dec_f = timed(dec_fused, iters)
    ok = ok and bool(torch.equal(out, m))
    # single-item latency (batch = 1)
    kg1 = timed(lambda: bfv.keygen(sk, pk, batch=1), 50)
    enc1 = timed(lambda: bfv.encrypt(c, pk, m, batch=1), 50)
    dec1 = timed(lambda: bfv.decrypt(out, c, sk, batch=1), 50)
    # the same single-item calls replayed from CUDA graphs (no per-launch host cost)
    def graph_us(fn):
        st = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(st):
            fn(st)
            st.synchronize()
            with torch.cuda.graph(g, stream=st):
                fn(st)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 10.0

    gk = graph_us(lambda st: bfv.keygen(sk, pk, batch=1, stream=st))
    ge = graph_us(lambda st: bfv.encrypt(c, pk, m, batch=1, stream=st))
    gd = graph_us(lambda st: bfv.decrypt(out, c, sk, batch=1, stream=st))
    gef = graph_us(lambda st: bfv.encrypt(c, None, m, batch=1, stream=st))
    gdf = graph_us(lambda st: bfv.decrypt(out, c, None, batch=1, stream=st))
    res = {"set": setname, "n": n, "limbs": r, "batch": batch, "roundtrip_ok": ok,
           "keygen_per_s": batch / (kg * 1e-3), "encrypt_per_s": batch / (enc * 1e-3), "decrypt_per_s": batch / ((dec_ms - cp_ms) * 1e-3),
           "enc_plus_dec_per_s": batch / ((enc + dec_ms - cp_ms) * 1e-3),
           "fused": {"encrypt_per_s": batch / (enc_f * 1e-3), "decrypt_per_s": batch / ((dec_f - cp_ms) * 1e-3),
                     "enc_plus_dec_per_s": batch / ((enc_f + dec_f - cp_ms) * 1e-3), "encrypt_ms": enc_f, "decrypt_ms": dec_f - cp_ms},
           "keygen_ms": kg, "encrypt_ms": enc, "decrypt_ms": dec_ms - cp_ms,
           "single_item_us": {"keygen": kg1 * 1e3, "encrypt": enc1 * 1e3, "decrypt": dec1 * 1e3},
           "single_item_graph_us": {"keygen": gk, "encrypt": ge, "decrypt": gd, "encrypt_fused": gef, "decrypt_fused": gdf}}
    bfv.close()
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="32k_16q")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    print(json.dumps(run(a.set, a.batch, a.iters)))
