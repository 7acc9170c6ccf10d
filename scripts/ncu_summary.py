"""Summarises an ncu report (--set full) into a markdown table + a small JSON with the numbers bench.py quotes.
   python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_final   ->  profiles/r01_final_ncu_full_summary.md, _pipe_util.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__issue_active.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    md = ["# ncu --set full --clock-control none summary (%s)" % rep, "", note, ""]
    js = {}
    seen = set()
    for r in rows[2:]:
        name = r[ki]
        if name in seen:
            continue
        seen.add(name)
        md += ["## " + name, "", "| metric | value | unit |", "|---|---|---|"]
        d = {}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                md.append("| %s | %s | %s |" % (k, r[i], units[i]))
                try:
                    d[k] = float(r[i].replace(",", ""))
                except ValueError:
                    d[k] = r[i]
        if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d:
            ur = units[hdr.index("dram__bytes_read.sum")]
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(ur, 1.0)
            d["traffic_bytes"] = (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) * scale
            md.append("| **traffic = dram read + write** | %.1f | Mbyte |" % (d["traffic_bytes"] / 1e6))
        md.append("")
        js[name] = d
    open(prefix + "_ncu_full_summary.md", "w").write("\n".join(md))
    json.dump(js, open(prefix + "_pipe_util.json", "w"), indent=1)
    print("wrote", prefix + "_ncu_full_summary.md", "and", prefix + "_pipe_util.json", "for", len(js), "kernels")


if __name__ == "__main__":
    main()
