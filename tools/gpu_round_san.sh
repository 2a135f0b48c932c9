#!/bin/bash
# compute-sanitizer memcheck over a subset of the parity tests (new kernels: LR, pairs, CSR prologue, SpMM T=4)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_interfaces.py tests/test_gpu_parity.py -m gpu -x -q \
  -k "leftright or matrix_free or shared_columns or edge_shapes or batched_selld or operator_mul or config1 or tls or expval_fused or uniform_width or tfim" \
  > gpurun_out/san_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/san_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/san_memcheck.log; tail -12 gpurun_out/san_memcheck.log
