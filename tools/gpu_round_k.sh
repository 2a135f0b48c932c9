#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/k_pytest.log
tail -3 gpurun_out/k_pytest.log
timeout 600 python tools/bench_configs.py --configs 4 --liou-spins 10 --newton-steps 10 >> gpurun_out/k_configs.jsonl 2>> gpurun_out/k_configs.err
cut -c1-400 gpurun_out/k_configs.jsonl; tail -5 gpurun_out/k_configs.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k_launches_newton.csv python tools/bench_configs.py --configs 4 --liou-spins 10 --newton-steps 2 > gpurun_out/k_launches.log 2>&1
