#!/bin/bash
TAG=${1:-var}
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for lib in libnttb200.so $(cd ntt-cuda_b200/nttb200 && ls libnttb200_*.so); do
  NTTB200_LIB=$PWD/ntt-cuda_b200/nttb200/$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_$lib.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_$lib.json").read().strip().splitlines()[-1])
print("$lib value %.4g ms/step %.4f"%(d["value"], d["ms_per_step"]), d["kernels_ms"], "inv %.4g"%d["inverse"]["value"])
PY
done
