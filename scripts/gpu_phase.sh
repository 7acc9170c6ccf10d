#!/bin/bash
mkdir -p gpurun_out
for lib in libnttb200.so $(cd ntt-cuda_b200/nttb200 && ls libnttb200_*.so 2>/dev/null); do
  echo "== $lib"; NTTB200_LIB=$PWD/ntt-cuda_b200/nttb200/$lib timeout 300 python scripts/phase_probe.py 2>&1 | tail -1 | tee gpurun_out/phase_$lib.json
done
