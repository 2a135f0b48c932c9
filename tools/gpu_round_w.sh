#!/bin/bash
# GPU session W2: ncu capture of the final headline kernel + launch list of bench.py.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_spmv_selld -s 30 -c 2 -o gpurun_out/w_selld python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/w_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/w_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/w_launches.log 2>&1
ls -la gpurun_out | grep "w_"
