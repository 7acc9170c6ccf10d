"""CPU-side checks of the C ABI: the library loads and exports every symbol include/nttb200.h declares."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import nttb200
    hdr = open(os.path.join(ROOT, "include", "nttb200.h")).read()
    names = set(re.findall(r"NTTB200_API\s+[\w\s\*]+?\b(nttb200_\w+)\s*\(", hdr))
    assert len(names) >= 10
    lib = nttb200.lib()
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.nttb200_version() >= 100
    assert b"invalid" in lib.nttb200_error_string(10001)


def test_no_oracle_in_product_path():
    """The product must never import / link the oracle (tier rule 3)."""
    pkg = os.path.join(ROOT, "ntt-cuda_b200")
    for root, _, files in os.walk(pkg):
        if os.sep + "build" in root:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in txt and "ntt_oracle" not in txt and "from oracle" not in txt, f
