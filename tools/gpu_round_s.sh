#!/bin/bash
# GPU session S: launch list of the config-4 Newton step (N = 2^22) -- kernel shares.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s_newton_launches.csv \
  python tools/bench_configs.py --configs 4 --liou-spins 11 --newton-steps 2 > gpurun_out/s_newton.log 2>&1
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open('gpurun_out/s_newton_launches.csv')) if len(r) > 14 and r[0].isdigit()]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = re.sub(r'\(.*', '', r[4]); tot[name] += float(r[14]); cnt[name] += 1
s = sum(tot.values())
for k, v in tot.most_common(25): print(f'{k[:70]:70s} n={cnt[k]:5d} total {v/1e6:9.3f} ms  {100*v/s:5.1f}%  avg {v/cnt[k]/1e3:8.1f} us')
PY
tail -3 gpurun_out/s_newton.log | cut -c1-400
