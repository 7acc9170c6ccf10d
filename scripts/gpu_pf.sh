#!/bin/bash
# L2-prefetch distance sweep + build variants; prints one line per configuration.
TAG=${1:-pf}
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.txt
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bfv 2>&1 | tail -1 > gpurun_out/${TAG}_$name.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$name.json").read().strip().splitlines()[-1])
    print("$name value %.4g ms/step %.4f"%(d["value"], d["ms_per_step"]), d.get("kernels_ms"), "inv %.4g"%d["inverse"]["value"], d["inverse"].get("kernels_ms"))
except Exception as e: print("$name FAILED", e, open("gpurun_out/${TAG}_$name.json").read()[-300:])
PY
}
for w in ${PF_LIST:-0 0.5 1 2}; do run pf$w NTTB200_PF_WAVES=$w; done
for lib in $(cd ntt-cuda_b200/nttb200 && ls libnttb200_*.so 2>/dev/null); do run $lib NTTB200_LIB=$PWD/ntt-cuda_b200/nttb200/$lib; done
