"""Markdown table from an ncu launch list (CSV written by `ncu --csv --log-file ... --metrics
gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread`): one row per launch, totals at the end.
Usage: python scripts/launch_table.py file.csv [first_id [last_id]]"""
import csv
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    h = rows[0]
    idc, kc, mc, vc = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
    launches = OrderedDict()
    for r in rows[1:]:
        i = int(r[idc])
        if lo <= i <= hi:
            launches.setdefault(i, {"name": r[kc]})[r[mc]] = float(r[vc].replace(",", ""))
    print("| # | kernel | us | DRAM read MB | DRAM write MB | regs |\n|---|---|---|---|---|---|")
    tot = [0.0, 0.0, 0.0]
    for i, L in launches.items():
        name = L["name"].split("(")[0].replace("void ", "")
        if "<" in L["name"]:
            name = L["name"].replace("void ", "").split(">(")[0] + ">"
        us, rd, wr = L.get("gpu__time_duration.sum", 0) / 1e3, L.get("dram__bytes_read.sum", 0) / 1e6, L.get("dram__bytes_write.sum", 0) / 1e6
        tot = [tot[0] + us, tot[1] + rd, tot[2] + wr]
        print("| %d | `%s` | %.1f | %.1f | %.1f | %d |" % (i, name, us, rd, wr, int(L.get("launch__registers_per_thread", 0))))
    print("| | **total** | **%.1f** | **%.1f** | **%.1f** | |" % tuple(tot))


if __name__ == "__main__":
    main()
