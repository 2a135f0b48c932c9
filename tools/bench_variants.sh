#!/bin/bash
# Runs bench.py (device-resident timing only) for each SELL kernel variant; prints one line each.
mkdir -p gpurun_out
run() {
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" 2>>gpurun_out/variants.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
r = d['roofline']
print('%-28s %8.1f steps/s  %7.1f us/launch  %7.1f GB/s  frac %.3f  e2e %.1f' % ('$TAG', d['value'], r['avg_launch_us'], r['achieved'], r['frac'], d['e2e']['value']))
"
}
if [ -z "$CFGS" ]; then
  TAG="csr"          run --format csr
  TAG="sell/ldg"     QPROP_SELL_KERNEL=ldg run --format sell
fi
for c in ${CFGS:-0 1 2 3 4 5}; do
  TAG="sell/tma cfg$c" QPROP_SELL_KERNEL=tma QPROP_TMA_CFG=$c run --format sell
done
