#!/bin/bash
# Register-only butterfly micro-benchmark + ncu pipe/stall sections for selected variants.
mkdir -p gpurun_out
B=ntt-cuda_b200/build/bfly_ubench
$B > gpurun_out/bfly_ubench.json 2> gpurun_out/bfly_ubench.err
cat gpurun_out/bfly_ubench.json
for v in ${NCU_VARIANTS:-fwd_lazy_approx_3cta fwd_lazy_approx_regtw_3cta fwd_lazy_v3_regtw_2cta}; do
  ncu --clock-control none --section SchedulerStats --section WarpStateStats --section ComputeWorkloadAnalysis --section InstructionStats --section Occupancy \
      -s 1 -c 1 --csv --page raw --log-file gpurun_out/ncu_ubench_$v.csv $B $v > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/ncu_ubench_$v.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
if hdr:
    h=rows[hdr[0]]; val=rows[hdr[0]+2]
    keep=("smsp__issue_active.avg.per_cycle_active","smsp__inst_executed.sum","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active","smsp__warps_active.avg.per_cycle_active","smsp__warps_eligible.avg.per_cycle_active","launch__registers_per_thread")
    print("$v")
    for k,x in zip(h,val):
        if k in keep or "issue_stalled" in k and "per_warp_active" in k or "pipe_" in k and "pct_of_peak_sustained_active" in k: print("   ",k,x)
PY
done
