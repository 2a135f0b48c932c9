#!/bin/bash
# GPU session T3: PDL in the SELL-D kernel (config 2 bench) + parity.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/t_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_pytest.log
tail -6 gpurun_out/t_pytest.log
for p in 0 1 0 1; do
QPROP_PDL=$p python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>gpurun_out/t.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']
print('pdl=$p  %8.1f steps/s  %7.2f us/launch  frac_stored %.3f  e2e %.1f normdev %.2e' % (d['value'], r['avg_launch_us'], r['frac_stored'], d['e2e']['value'], d['config']['norm_deviation_after_run']))"
done
tail -3 gpurun_out/t.err
