#!/bin/bash
# GPU session A: parity tests, bench variants of the SELL-D kernel, launch list + ncu capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
run() {
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" 2>>gpurun_out/a_variants.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
r = d['roofline']
print('%-28s %8.1f steps/s  %7.1f us/launch  alg %7.1f GB/s frac %.3f  stored %7.1f GB/s frac %.3f  e2e %.1f normdev %.2e' % ('$TAG', d['value'], r['avg_launch_us'], r['achieved'], r['frac'], r['achieved_stored'], r['frac_stored'], d['e2e']['value'], d['config']['norm_deviation_after_run']))
"
}
{
TAG="sell/tma"       run --format sell
for mb in 2 3 4; do
  TAG="selld minb$mb" QPROP_SELLD_MINB=$mb run --format selld
done
for c in 1 2 4 6 8; do
  TAG="selld minb2 ctas$c" QPROP_SELLD_MINB=2 QPROP_SELLD_CTAS=$c run --format selld
done
TAG="selld minb4 ctas8" QPROP_SELLD_MINB=4 QPROP_SELLD_CTAS=8 run --format selld
} > gpurun_out/a_variants.txt 2>&1
cat gpurun_out/a_variants.txt
python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
cat gpurun_out/a_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/a_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spmv_selld -s 30 -c 2 -o gpurun_out/a_selld python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/a_ncu.log 2>&1
ls -la gpurun_out
