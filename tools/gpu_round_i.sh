#!/bin/bash
# profiling session: launch list + ncu --set full of the three dominant kernels with the default settings
mkdir -p gpurun_out
python bench.py > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; cut -c1-300 gpurun_out/i_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/i_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spmv_selld -s 30 -c 2 -o gpurun_out/i_selld python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/i_ncu_selld.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_selld -s 30 -c 1 -o gpurun_out/i_spmm python tools/bench_configs.py --configs 3 --B 1024 --steps 2 > gpurun_out/i_ncu_spmm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_dense -s 4 -c 1 -o gpurun_out/i_gemm python tools/bench_configs.py --configs 5 --dense-B 64 > gpurun_out/i_ncu_gemm.log 2>&1
./tools/micro/dmma_peak > gpurun_out/i_dmma_peak.jsonl 2>&1
timeout 900 python tools/bench_configs.py --configs 1,3,4,5 --B 1024 > gpurun_out/i_configs.jsonl 2> gpurun_out/i_configs.err
python tools/calibrate_fp64.py > gpurun_out/i_fp64_peak.json 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/i_bench_ref.json 2> gpurun_out/i_bench_ref.err
python bench.py --format sell --no-cpu-baseline > gpurun_out/i_bench_sell.json 2>/dev/null
tail -3 gpurun_out/i_configs.err
