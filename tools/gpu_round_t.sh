#!/bin/bash
# GPU session T4: PDL in the Krylov kernels -- parity + config 4 (n = 10, 11) with and without.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/t_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_pytest.log
tail -6 gpurun_out/t_pytest.log
for p in 0 1; do for n in 10 11; do
QPROP_PDL=$p timeout 600 python tools/bench_configs.py --configs 4 --liou-spins $n --newton-steps 5 2>>gpurun_out/t.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('pdl=$p n=$n  ms/step %.3f' % d['ms_per_step'])"
done; done
tail -3 gpurun_out/t.err
