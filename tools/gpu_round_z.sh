#!/bin/bash
# GPU session Z: matrix-free Liouvillian kernel -- parity + config 4 (Newton) timings.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_interfaces.py -m gpu -x -q ) > gpurun_out/z_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/z_pytest.log
tail -6 gpurun_out/z_pytest.log
rm -f gpurun_out/z_configs.jsonl
for n in 10 11 12; do
  timeout 900 python tools/bench_configs.py --configs 4 --liou-spins $n --newton-steps 3 --matrix-free >> gpurun_out/z_configs.jsonl 2>> gpurun_out/z.err
done
python - <<'PY'
import json
for l in open('gpurun_out/z_configs.jsonl'):
    d = json.loads(l); print(d['workload'][:60], 'ms/step %.2f' % d['ms_per_step'], 'mem %.2f GB' % d['device_memory_used_gb'])
PY
tail -5 gpurun_out/z.err
