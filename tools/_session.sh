mkdir -p gpurun_out
# launch list of bench.py (kernel shares of the step): every launch of the library's kernels, device time each
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --min-seconds 0.01 --no-cpu-baseline > gpurun_out/r2_launches.log 2>&1
tail -2 gpurun_out/r2_launches.log | cut -c1-200
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2_launches.csv')) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split('(')[0][:70]; val = float(r[-1].replace(',', ''))
    agg[name][0] += 1; agg[name][1] += val
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]: print(f'{t:14.1f} {n:5d}  {k}')
PY
# full capture of the headline kernel in the same command
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_selld -s 30 -c 1 -o gpurun_out/r2_selld python bench.py --steps 2 --warmup 1 --min-seconds 0.01 --no-cpu-baseline --no-ensemble --no-sell-arm > gpurun_out/r2_selld_ncu.log 2>&1; tail -1 gpurun_out/r2_selld_ncu.log
