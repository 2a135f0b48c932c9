#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "dense or dmma" > gpurun_out/j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/j_pytest.log
tail -3 gpurun_out/j_pytest.log
for v in 0 1 2 3; do
QPROP_DMMA_VARIANT=$v timeout 600 python tools/bench_configs.py --configs 5 --dense-B 64 2>> gpurun_out/j_configs.err | sed "s/^{/{\"dmma_variant\": $v, /" >> gpurun_out/j_configs.jsonl
done
timeout 600 python tools/bench_configs.py --configs 5 --dense-B 16,32,128 >> gpurun_out/j_configs.jsonl 2>> gpurun_out/j_configs.err
cut -c1-330 gpurun_out/j_configs.jsonl; tail -5 gpurun_out/j_configs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_dense -s 4 -c 1 -o gpurun_out/j_gemm python tools/bench_configs.py --configs 5 --dense-B 64 > gpurun_out/j_ncu_gemm.log 2>&1
