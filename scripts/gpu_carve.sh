#!/bin/bash
for c in default 45 50 60 75 100; do
  if [ $c = default ]; then unset NTTB200_CARVEOUT; else export NTTB200_CARVEOUT=$c; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bfv | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('carveout $c', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernels_ms'].items()}, 'inv', round(d['inverse']['ms_per_step'],4))"
done
