#!/bin/bash
# GPU session T5: contiguous ranges vs interleaved chunks in the SELL-D kernel (config 2).
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "selld or tfim or fullsize" ) > gpurun_out/t_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_pytest.log
tail -3 gpurun_out/t_pytest.log
for c in 0 8 16 32 64; do
QPROP_SELLD_INTER=$c python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>gpurun_out/t.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); r = d['roofline']
print('inter=$c  %8.1f steps/s  %7.2f us/launch  frac_stored %.3f  normdev %.2e' % (d['value'], r['avg_launch_us'], r['frac_stored'], d['config']['norm_deviation_after_run']))"
done
tail -3 gpurun_out/t.err
