#!/bin/bash
# tests + bench sweep over the persistent-kernel iteration target
TAG=${1:-sweep}
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.txt
for it in 1 2 4 8 16 64; do
  echo "=== ITERS=$it"
  NTTB200_ITERS=$it timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${TAG}_iters$it.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_iters$it.json").read().strip().splitlines()[-1])
print("ITERS=$it value %.4g ms/step %.4f"%(d["value"], d["ms_per_step"]), d["kernels_ms"], "inv %.4g"%d["inverse"]["value"])
PY
done
