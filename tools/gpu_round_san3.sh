#!/bin/bash
# compute-sanitizer memcheck: dense (DMMA) path, expectation values, small full-format checks
mkdir -p gpurun_out
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_ensemble.py -m gpu -x -q \
  -k "dense or expval_dense or state_verbs or host_operator_verbs or explicit_diagonal or timings or reinit or storage_and_parameters" \
  > gpurun_out/san3_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/san3_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/san3_memcheck.log; tail -6 gpurun_out/san3_memcheck.log
