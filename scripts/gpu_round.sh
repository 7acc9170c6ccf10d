#!/bin/bash
# One fat gpurun call: parity tests, integer-pipe peaks, bench, ncu launch list.  Usage: gpurun -- bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.txt
echo "=== ubench"; timeout 120 ./ntt-cuda_b200/build/ipipe_ubench | tee gpurun_out/${TAG}_ipipe.json
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -3 | tee gpurun_out/${TAG}_bench.json
echo "=== bench no-tma"; NTTB200_NO_TMA=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_notma.json
echo "=== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_bench.log
