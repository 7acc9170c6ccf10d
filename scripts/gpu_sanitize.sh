#!/bin/bash
# compute-sanitizer over a small slice of the GPU suite (memcheck + racecheck + synccheck); summaries into gpurun_out/
mkdir -p gpurun_out
SEL='test_ctx_forward_inverse_vs_oracle and (11-1-3 or 13-1-2 or 15-16-40) or test_fused_polynomial_product_vs_oracle and (12-3-4 or 14-5-6) or test_pipelines_vs_oracle and 8k_4q or test_loaded_key_fused_path and (57-13-4) or test_homomorphic_add_and_plain_multiply and 4k_3q or test_wire_format_round_trip and 4k_3q'
for tool in memcheck racecheck synccheck; do
  echo "=== $tool"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_bfv.py -x -q -m gpu -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitize_$tool.log | tail -5
done
