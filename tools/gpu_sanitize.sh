#!/bin/bash
# compute-sanitizer over the GPU parity tests:  bash tools/gpu_sanitize.sh <memcheck|racecheck|synccheck> [pytest -k expression]
tool=${1:-memcheck}; sel=${2:-""}
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_interfaces.py -m gpu -x -q ${sel:+-k "$sel"} \
  > gpurun_out/san_${tool}.log 2>&1; echo "$tool exit $?" >> gpurun_out/san_${tool}.log
grep -c "Invalid\|out of bounds\|misaligned\|hazard\|Race reported\|Barrier error" gpurun_out/san_${tool}.log; tail -12 gpurun_out/san_${tool}.log
