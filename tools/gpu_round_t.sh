#!/bin/bash
# GPU session T7: L1::no_allocate for the gathers that leave the CTA's row range (config 2).
mkdir -p gpurun_out
for s in 0 1 0.25 4 0.0625; do
QPROP_SELLD_NEAR=$s python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>gpurun_out/t.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']
print('near=$s  %8.1f steps/s  %7.2f us/launch  frac_stored %.3f  normdev %.2e' % (d['value'], r['avg_launch_us'], r['frac_stored'], d['config']['norm_deviation_after_run']))"
done
tail -3 gpurun_out/t.err
