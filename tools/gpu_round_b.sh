#!/bin/bash
# GPU session B: parity tests, SELL-D variants, secondary configs.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b_pytest.log
tail -5 gpurun_out/b_pytest.log
run() {
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" 2>>gpurun_out/b_variants.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
r = d['roofline']
print('%-28s %8.1f steps/s  %7.1f us/launch  alg %7.1f GB/s frac %.3f  stored %7.1f GB/s frac %.3f  e2e %.1f normdev %.2e' % ('$TAG', d['value'], r['avg_launch_us'], r['achieved'], r['frac'], r['achieved_stored'], r['frac_stored'], d['e2e']['value'], d['config']['norm_deviation_after_run']))
"
}
{
for mb in 2 3 4; do
  TAG="selld minb$mb" QPROP_SELLD_MINB=$mb run --format selld
done
TAG="selld wide minb2" QPROP_SELLD_WIDE=1 QPROP_SELLD_MINB=2 run --format selld
TAG="selld wide minb3" QPROP_SELLD_WIDE=1 QPROP_SELLD_MINB=3 run --format selld
for spc in 16 32 64 128; do
  TAG="selld minb3 spc$spc" QPROP_SELLD_MINB=3 QPROP_SELLD_SPC=$spc run --format selld
done
TAG="selld minb2 spc64" QPROP_SELLD_MINB=2 QPROP_SELLD_SPC=64 run --format selld
TAG="selld minb2 ctas2 spc -" QPROP_SELLD_MINB=2 QPROP_SELLD_CTAS=2 run --format selld
} > gpurun_out/b_variants.txt 2>&1
cat gpurun_out/b_variants.txt
timeout 900 python tools/bench_configs.py --configs 3,4 --B 1024 > gpurun_out/b_configs.jsonl 2> gpurun_out/b_configs.err
cat gpurun_out/b_configs.jsonl; tail -5 gpurun_out/b_configs.err
