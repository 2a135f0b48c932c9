#!/bin/bash
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/n_mem.txt; nproc >> gpurun_out/n_mem.txt
( time timeout 900 python tools/bench_configs.py --configs 4 --liou-spins 12 --newton-steps 3 ) >> gpurun_out/n_configs.jsonl 2>> gpurun_out/n_configs.err
cut -c1-400 gpurun_out/n_configs.jsonl; tail -8 gpurun_out/n_configs.err; cat gpurun_out/n_mem.txt
