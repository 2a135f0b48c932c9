#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/d_pytest.log
tail -5 gpurun_out/d_pytest.log
timeout 900 python tools/bench_configs.py --configs 3,5,4 --B 1024 --dense-B 16,64 > gpurun_out/d_configs.jsonl 2> gpurun_out/d_configs.err
cat gpurun_out/d_configs.jsonl; tail -5 gpurun_out/d_configs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_selld -s 30 -c 1 -o gpurun_out/d_spmm python tools/bench_configs.py --configs 3 --B 1024 --steps 2 > gpurun_out/d_ncu_spmm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_dense -s 4 -c 1 -o gpurun_out/d_gemm python tools/bench_configs.py --configs 5 --dense-B 64 > gpurun_out/d_ncu_gemm.log 2>&1
ls -la gpurun_out | grep " d_"
