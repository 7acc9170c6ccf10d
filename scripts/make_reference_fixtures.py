"""Golden fixtures produced BY THE REFERENCE ITSELF: runs oracle/_ref/ref_dump (the unmodified reference headers + a dump driver,
rebuilt for sm_100a by oracle/ref/Makefile) on a B200 and records SHA-256 digests of every buffer the reference's own kernels
produce -- raw forward / inverse NTT outputs, Salsa20 keystream, secret key, public key, ciphertext (padding included), plaintext
-- plus the gaussian draws themselves (the one value a CPU cannot reproduce bit for bit: normcdfinvf), so that the CPU oracle can
be pinned against the reference offline (tests/test_oracle_golden.py::test_oracle_matches_reference_gpu_outputs).

    gpurun -- python scripts/make_reference_fixtures.py gpurun_out/reference_gpu_fixtures      (then copy into tests/golden/)

Inputs are seeded exactly as the tests seed them (splitmix64, oracle.fill_uniform): polynomial p of the NTT dump has seed
0x5EED0000 + p, the BFV message has seed 0xC0FFEE.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ntt-cuda_b200"))
from nttb200 import params  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "ref_dump")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out_prefix = sys.argv[1]
    fx, draws = {"generator": "scripts/make_reference_fixtures.py", "ntt": {}, "bfv": {}}, {}
    for name, num in (("4k_3q", 6), ("8k_3q", 7), ("32k_16q", 32)):
        with tempfile.TemporaryDirectory() as d:
            subprocess.run([EXE, "ntt", name, str(num), d], check=True, capture_output=True)
            fwd = np.fromfile(os.path.join(d, "ref_fwd.bin"), dtype=np.uint64)
            inv = np.fromfile(os.path.join(d, "ref_inv.bin"), dtype=np.uint64)
        n = params.RNS_SETS[name][0]
        fx["ntt"][name] = {"num": num, "fwd_sha256": sha(fwd), "inv_sha256": sha(inv),
                           "fwd_sha256_per_poly": [sha(fwd[p * n:(p + 1) * n]) for p in range(num)], "fwd_head": [int(v) for v in fwd[:8]]}
    for name in ("4k_3q", "8k_4q", "16k_5q"):
        n, qs, _ = params.RNS_SETS[name]
        r = len(qs)
        with tempfile.TemporaryDirectory() as d:
            res = subprocess.run([EXE, "bfv", name, d], check=True, capture_output=True, text=True)
            assert "roundtrip ok" in res.stdout
            ref = {k: np.fromfile(os.path.join(d, f"ref_{k}.bin"), dtype=np.uint8) for k in ("keygen_in", "sk", "pk", "temp", "c", "e", "plain")}
        u = lambda k: ref[k].view(np.uint64)
        q0 = int(qs[0])
        dec = lambda v: np.where(v > q0 // 2, v.astype(np.int64) - q0, v.astype(np.int64)).astype(np.int8)
        draws[f"{name}_keygen_e"] = dec(u("temp")[:n])                     # limb 0 of the residues = the signed draw
        draws[f"{name}_enc_e0"] = dec(u("e")[:n])
        draws[f"{name}_enc_e1"] = dec(u("e")[r * n:r * n + n])
        fx["bfv"][name] = {k + "_sha256": sha(ref[k]) for k in ref}
        fx["bfv"][name]["sk_head"] = [int(v) for v in u("sk")[:4]]
        fx["bfv"][name]["c_head"] = [int(v) for v in u("c")[:4]]
    json.dump(fx, open(out_prefix + ".json", "w"), indent=1)
    np.savez_compressed(out_prefix + "_draws.npz", **draws)
    print("wrote", out_prefix + ".json", out_prefix + "_draws.npz")


if __name__ == "__main__":
    main()
