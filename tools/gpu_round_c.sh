#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c_pytest.log
tail -5 gpurun_out/c_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; cat gpurun_out/c_bench.json
python tools/calibrate_fp64.py > gpurun_out/c_fp64_peak.json 2>&1; cat gpurun_out/c_fp64_peak.json
timeout 900 python tools/bench_configs.py --configs 3,4,5 --B 1024 --dense-B 16,64 > gpurun_out/c_configs.jsonl 2> gpurun_out/c_configs.err
cat gpurun_out/c_configs.jsonl; tail -5 gpurun_out/c_configs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_dense -s 4 -c 1 -o gpurun_out/c_gemm python tools/bench_configs.py --configs 5 --dense-B 64 > gpurun_out/c_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_selld -s 30 -c 1 -o gpurun_out/c_spmm python tools/bench_configs.py --configs 3 --B 1024 --steps 2 > gpurun_out/c_ncu_spmm.log 2>&1
ls -la gpurun_out | tail -12
