#!/bin/bash
# GPU session T6: CTA shape of the SELL-D kernel (config 2): 2 x 256 threads vs 1 x 512 vs 4 x 128 per SM.
mkdir -p gpurun_out
for t in 256 512 128 1024; do
QPROP_SELLD_THREADS=$t python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>gpurun_out/t.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']
print('threads=$t  %8.1f steps/s  %7.2f us/launch  frac_stored %.3f  normdev %.2e' % (d['value'], r['avg_launch_us'], r['frac_stored'], d['config']['norm_deviation_after_run']))"
done
tail -3 gpurun_out/t.err
