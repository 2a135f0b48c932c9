#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -2
for t in 0 1 0 1; do
QPROP_SELLD_REAL=$t python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>gpurun_out/t.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']
print('real=$t %8.1f steps/s  %7.2f us/launch  frac_stored %.3f  e2e %.1f normdev %.2e' % (d['value'], r['avg_launch_us'], r['frac_stored'], d['e2e']['value'], d['config']['norm_deviation_after_run']))"
done
timeout 600 python tools/bench_configs.py --configs 4 --liou-spins 11 --newton-steps 3 2>>gpurun_out/t.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('config 4 n=11 ms/step %.3f' % d['ms_per_step'])"
tail -3 gpurun_out/t.err
