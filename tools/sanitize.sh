#!/bin/bash
# compute-sanitizer over the kernel-level and step-level GPU tests: memcheck, racecheck, synccheck.
# Run on a GPU box:  bash tools/sanitize.sh  -> gpurun_out/r2_sanitizer_*.log ; summaries go to profiles/.
set -u
mkdir -p gpurun_out
TESTS="tests/test_gpu_kernels.py tests/test_gpu_chain.py::test_chain_is_bit_identical_to_per_kernel_orchestration"
for tool in memcheck racecheck synccheck; do
  timeout 480 compute-sanitizer --tool $tool --error-exitcode 0 --print-limit 20 \
    python -m pytest $TESTS -m gpu -q -x -k "not full_size and not cfg_kw3" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?" >> gpurun_out/r2_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/r2_sanitizer_$tool.log | tail -5
done
