#!/bin/bash
# GPU session W3: ncu captures of the final kernels (SELL-D real-table/tail variant, pair kernel) + launch list.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_spmv_selld -s 30 -c 1 -o gpurun_out/w_selld python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/w_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm_selld_pairs -s 30 -c 1 -o gpurun_out/w_pairs python tools/bench_configs.py --configs 3 --B 1024 --steps 2 > gpurun_out/w_ncu_pairs.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/w_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/w_launches.log 2>&1
ls -la gpurun_out | grep "w_"
