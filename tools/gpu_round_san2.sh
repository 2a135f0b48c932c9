#!/bin/bash
# compute-sanitizer memcheck over the Krylov / dense / batched / propagator tests
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "newton or arnoldi or specrange or batched_per_trajectory or protocol or one_call or norms_batched or check_normalization" \
  > gpurun_out/san2_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/san2_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/san2_memcheck.log; tail -6 gpurun_out/san2_memcheck.log
