"""Static SASS instruction-class counts of the n = 2^15 NTT kernels (they are straight-line code per thread, so the static count is
the per-thread dynamic count up to the uniform branches) and the fma-heavy pipe cycles per thread they imply under the measured
issue model (profiles/r01_ipipe2.md: IMAD.WIDE / IMAD.HI 4 cycles per warp instruction, other IMAD forms 2).
   python ntt-cuda_b200/tools/sass_pipe_count.py > profiles/r01_sass_counts.json"""
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OBJ = os.path.join(HERE, "..", "build", "ntt_launch.o")
KERNELS = {
    "ntt_strided_pass_fwd": "16ntt_strided_passINS_15ShoupLazyPolicyELi15ELb0",
    "ntt_contig_pass_fwd": "15ntt_contig_passINS_15ShoupLazyPolicyELi15ELb0",
    "ntt_contig_pass_inv": "15ntt_contig_passINS_18ShoupLazyInvPolicyELi15ELb1",
    "ntt_strided_pass_inv": "16ntt_strided_passINS_18ShoupLazyInvPolicyELi15ELb1",
}
BUTTERFLIES_PER_THREAD = {"ntt_strided_pass_fwd": 64, "ntt_strided_pass_inv": 64, "ntt_contig_pass_fwd": 56, "ntt_contig_pass_inv": 56}
THREADS_PER_LAUNCH = {k: 1024 * 32768 // 16 for k in KERNELS}     # C2: 1024 polynomials, 16 coefficients per thread


def main():
    sass = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    out = {}
    for name, mangled in KERNELS.items():
        body, on = [], False
        for line in sass.splitlines():
            if "Function :" in line:
                on = mangled in line
            elif on:
                m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_.]+)", line)
                if m:
                    body.append(m.group(1))
        c = {"IMAD.WIDE": 0, "IMAD.HI": 0, "IMAD_other": 0, "ALU": 0, "LSU": 0, "other": 0}
        for op in body:
            if op.startswith("IMAD.WIDE"):
                c["IMAD.WIDE"] += 1
            elif op.startswith("IMAD.HI"):
                c["IMAD.HI"] += 1
            elif op.startswith("IMAD") or op.startswith("HFMA2") or op.startswith("FFMA"):
                c["IMAD_other"] += 1
            elif re.match(r"(IADD3|LOP3|SHF|SEL|ISETP|MOV|PRMT|LEA|VIADD|PLOP3|IABS|FSEL)", op):
                c["ALU"] += 1
            elif re.match(r"(LDS|STS|LDG|STG|LDL|STL|LDC|CCTL)", op):
                c["LSU"] += 1
            else:
                c["other"] += 1
        fma_cycles = 4 * (c["IMAD.WIDE"] + c["IMAD.HI"]) + 2 * c["IMAD_other"]
        issue_model = fma_cycles + c["ALU"] + c["LSU"] + c["other"]
        b = BUTTERFLIES_PER_THREAD[name]
        out[name] = {"instructions": len(body), **c, "butterflies_per_thread": b,
                     "fma_heavy_cycles_per_warp": fma_cycles, "issue_model_cycles_per_warp": issue_model,
                     "fma_heavy_cycles_per_warp_butterfly": round(fma_cycles / b, 2),
                     "warps_per_launch_c2": THREADS_PER_LAUNCH[name] // 32}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
