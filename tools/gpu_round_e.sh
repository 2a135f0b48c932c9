#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/e_pytest.log
tail -4 gpurun_out/e_pytest.log
for t in 1 2; do
QPROP_SPMM_T=$t timeout 900 python tools/bench_configs.py --configs 3 --B 1024 >> gpurun_out/e_configs.jsonl 2>> gpurun_out/e_configs.err
done
timeout 900 python tools/bench_configs.py --configs 5 --dense-B 16,64 >> gpurun_out/e_configs.jsonl 2>> gpurun_out/e_configs.err
cat gpurun_out/e_configs.jsonl; tail -5 gpurun_out/e_configs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_dense -s 4 -c 1 -o gpurun_out/e_gemm python tools/bench_configs.py --configs 5 --dense-B 64 > gpurun_out/e_ncu_gemm.log 2>&1
QPROP_SPMM_T=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_selld -s 30 -c 1 -o gpurun_out/e_spmm python tools/bench_configs.py --configs 3 --B 1024 --steps 2 > gpurun_out/e_ncu_spmm.log 2>&1
