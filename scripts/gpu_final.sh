#!/bin/bash
# One fat gpurun call that refreshes every number quoted in DESIGN.md / profiles/.  Usage: gpurun -- bash scripts/gpu_final.sh [tag]
TAG=${1:-r01_final}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/${TAG}_pytest.txt
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench (ours)"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $O/${TAG}_bench.json; cut -c1-600 $O/${TAG}_bench.json
echo "=== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $O/${TAG}_bench_reference.json; cut -c1-400 $O/${TAG}_bench_reference.json
echo "=== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_ncu_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bfv > $O/${TAG}_ncu_launches.log 2>&1; tail -1 $O/${TAG}_ncu_launches.log | cut -c1-100
echo "=== ncu full (forward + inverse kernels)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 4 -c 2 -o $O/${TAG}_fwd python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bfv > $O/${TAG}_ncu_fwd.log 2>&1; tail -1 $O/${TAG}_ncu_fwd.log | cut -c1-100
# the inverse kernels are launches 2 and 3 of the ntt_ kernels (bench.py's round-trip gate runs forward then inverse first)
timeout 900 ncu --set full --clock-control none -k regex:ntt_ -s 2 -c 2 -o $O/${TAG}_inv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bfv > $O/${TAG}_ncu_inv.log 2>&1; tail -1 $O/${TAG}_ncu_inv.log | cut -c1-100
# summarise on the box and keep only the forward report (gpurun_out/ is capped at 64 MiB)
python scripts/ncu_summary.py $O/${TAG}_fwd.ncu-rep $O/${TAG} "forward kernels" | tail -1
python scripts/ncu_summary.py $O/${TAG}_inv.ncu-rep $O/${TAG}_inv "inverse kernels" | tail -1
rm -f $O/${TAG}_inv.ncu-rep
echo "=== phase probe"; timeout 300 python scripts/phase_probe.py 2>&1 | tail -1 | tee $O/${TAG}_phase_probe.json
echo "=== micro-benchmarks"; timeout 120 ./ntt-cuda_b200/build/ipipe2_ubench > $O/${TAG}_ipipe2_ubench.json; timeout 120 ./ntt-cuda_b200/build/bfly_ubench > $O/${TAG}_bfly_ubench.json; tail -3 $O/${TAG}_bfly_ubench.json | cut -c1-120
echo "=== bfv"; for s in 4k_3q 8k_4q 16k_5q 32k_9q 32k_16q; do timeout 300 python scripts/bfv_bench.py --set $s --batch 64 | tail -1; done > $O/${TAG}_bfv.jsonl; timeout 300 python scripts/bfv_bench.py --set 32k_16q --batch 256 | tail -1 >> $O/${TAG}_bfv.jsonl; timeout 300 python scripts/bfv_bench.py --set 8k_3q --batch 256 | tail -1 >> $O/${TAG}_bfv.jsonl; cut -c1-200 $O/${TAG}_bfv.jsonl
echo "=== polymul"; timeout 300 python scripts/polymul_bench.py | tail -1 | tee $O/${TAG}_polymul.json
echo "=== sweep"; timeout 600 python scripts/ntt_sweep.py > $O/${TAG}_ntt_sweep.jsonl 2>&1; tail -7 $O/${TAG}_ntt_sweep.jsonl | cut -c1-160
