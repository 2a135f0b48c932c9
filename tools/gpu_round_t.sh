#!/bin/bash
mkdir -p gpurun_out
for l in 0 24 48 96; do
QPROP_L2_PERSIST=$l python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>gpurun_out/t.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']
print('l2persist=$l MB  %8.1f steps/s  %7.2f us/launch  normdev %.2e' % (d['value'], r['avg_launch_us'], d['config']['norm_deviation_after_run']))"
done
tail -3 gpurun_out/t.err
