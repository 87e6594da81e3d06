#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over a slice of the parity suite that reaches every kernel family (eigen-form and
# matrix-form walks, tensor-pipe walks, fused incremental evaluation, device matrix combination; not the cross-device spin of reduce groups: the tool serialises launches); run on the GPU box.
mkdir -p gpurun_out
SEL='primates_golden[JC69] or tiny_test or oracle_parity[none-50-257-4-4] or oracle_parity[always-33-500-5-4] or oracle_parity[always-16-130-4-20] or oracle_parity[none-10-96-1-61] or by_partition or nan_signalling'
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_preorder_oracle.py tests/test_gpu_parity2.py -m gpu -q -x -k "$SEL or fused_incremental or mcmc_reenactment or walk_variants or matrix_convolution or preorder_matches_oracle_and_finite_differences[4-4-40-300] or preorder_matches_oracle_and_finite_differences[20-2-12-90] or cross_products_match_oracle[4-4-40-700] or cross_products_match_oracle[61-2-8-75] or plan_cache" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
