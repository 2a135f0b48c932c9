#!/bin/bash
# GPU session U2: pair kernel (operators sharing columns) -- parity + config 3.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/u_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/u_pytest.log
tail -12 gpurun_out/u_pytest.log
for np_ in 1 0; do
  if [ $np_ = 1 ]; then export QPROP_NO_PAIRS=1; else unset QPROP_NO_PAIRS; fi
  timeout 600 python tools/bench_configs.py --configs 3 --steps 3 2>>gpurun_out/u.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  no_pairs=$np_: us/term %.1f  frac %.3f  traj-steps/s %.0f normdev %.1e' % (d['us_per_term'], d['frac_of_measured_hbm'], d['trajectory_steps_per_s'], d['norm_dev_max']))"
done
tail -3 gpurun_out/u.err
