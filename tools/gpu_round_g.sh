#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/g_pytest.log
tail -4 gpurun_out/g_pytest.log
for v in 0 1 2 3; do
QPROP_DMMA_VARIANT=$v timeout 600 python tools/bench_configs.py --configs 5 --dense-B 64 2>> gpurun_out/g_configs.err | sed "s/^{/{\"dmma_variant\": $v, /" >> gpurun_out/g_configs.jsonl
done
python bench.py --no-cpu-baseline >> gpurun_out/g_configs.jsonl 2>> gpurun_out/g_configs.err
QPROP_SELLD_NO_CONST=1 python bench.py --no-cpu-baseline >> gpurun_out/g_configs.jsonl 2>> gpurun_out/g_configs.err
timeout 900 python tools/bench_configs.py --configs 1,3 --B 1024 >> gpurun_out/g_configs.jsonl 2>> gpurun_out/g_configs.err
cut -c1-420 gpurun_out/g_configs.jsonl; tail -5 gpurun_out/g_configs.err
ncu --set full --clock-control none --import-source on -k regex:k_spmv_selld -s 30 -c 1 -o gpurun_out/g_selld python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/g_ncu.log 2>&1
