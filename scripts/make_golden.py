"""Extracts the reference's only known-answer vector into tests/golden/decryption_kat.npz.

Source: /root/reference/BFV_Scheme/decryption_test.cu  (c_host :348, sk_host :355, parameter set :47-48,
expected plaintext m[i] = i % 10 :230-232).  Run in the build container only (the reference tree does not
exist on the GPU box); the .npz it writes is committed.
"""
import os
import re
import sys

import numpy as np

REF = "/root/reference/BFV_Scheme/decryption_test.cu"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "decryption_kat.npz")


def grab(text, name):
    m = re.search(r"^\s*unsigned long long %s\[\] = \{([0-9, ]*)\}" % name, text, re.M)
    if not m:
        sys.exit("cannot find %s in %s" % (name, REF))
    return np.array([int(v) for v in m.group(1).split(",") if v.strip()], dtype=np.uint64)


def main():
    text = open(REF).read()
    c_host = grab(text, "c_host")
    sk_host = grab(text, "sk_host")
    assert c_host.size == 24576 and sk_host.size == 8192, (c_host.size, sk_host.size)
    np.savez_compressed(OUT, c_host=c_host, sk_host=sk_host,
                        n=np.uint64(4096),
                        q=np.array([68719403009, 68719230977, 137438822401], dtype=np.uint64),
                        psi_roots=np.array([24250113, 29008497, 8625844], dtype=np.uint64),
                        t=np.uint64(1024), gamma=np.uint64(2305843009213683713))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
