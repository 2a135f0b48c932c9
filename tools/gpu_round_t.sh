#!/bin/bash
# GPU session T2: programmatic dependent launch in the CSR Chebyshev path -- parity + small-N scan.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/t_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_pytest.log
tail -6 gpurun_out/t_pytest.log
rm -f gpurun_out/t_scan.jsonl
for p in 0 1; do QPROP_PDL=$p timeout 600 python tools/micro/persist_scan.py 2>> gpurun_out/t_scan.err | sed "s/^{/{\"pdl\": $p, /" >> gpurun_out/t_scan.jsonl; done
python - <<'PY'
import json
for l in open('gpurun_out/t_scan.jsonl'):
    d = json.loads(l); print('pdl', d['pdl'], 'N', d['N'], 'n_coeffs', d['n_coeffs'], 'us/step %.1f' % d['us_per_step'], 'us/term %.2f' % d['us_per_term'], 'normdev %.1e' % d['norm_dev'])
PY
for p in 0 1; do QPROP_PDL=$p timeout 300 python tools/bench_configs.py --configs 1 2>>gpurun_out/t_scan.err | cut -c1-260; done
tail -3 gpurun_out/t_scan.err
