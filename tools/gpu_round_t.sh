#!/bin/bash
# GPU session T9: leaner SELL-D hot loop (interleaved table entry, 32-bit row + offset, FMA pairs).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "selld or tfim or fullsize" ) 2>&1 | tail -2
for i in 1 2; do
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>gpurun_out/t.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']
print('%8.1f steps/s  %7.2f us/launch  frac_stored %.3f  e2e %.1f normdev %.2e' % (d['value'], r['avg_launch_us'], r['frac_stored'], d['e2e']['value'], d['config']['norm_deviation_after_run']))"
done
timeout 600 python tools/bench_configs.py --configs 4 --liou-spins 11 --newton-steps 3 2>>gpurun_out/t.err | cut -c1-260
tail -3 gpurun_out/t.err
