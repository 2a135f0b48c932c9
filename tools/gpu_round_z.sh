#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_interfaces.py -m gpu -x -q ) 2>&1 | tail -2
for n in 11 12; do
  timeout 900 python tools/bench_configs.py --configs 4 --liou-spins $n --newton-steps 3 --matrix-free 2>> gpurun_out/z.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'][:60], 'ms/step %.2f' % d['ms_per_step'])"
done
