#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/f_pytest.log
tail -4 gpurun_out/f_pytest.log
./tools/micro/dmma_peak > gpurun_out/f_dmma_peak.jsonl 2>&1; cat gpurun_out/f_dmma_peak.jsonl | tr '\n' ' '; echo
timeout 900 python tools/bench_configs.py --configs 1,3 --B 1024 > gpurun_out/f_configs.jsonl 2> gpurun_out/f_configs.err
cat gpurun_out/f_configs.jsonl; tail -5 gpurun_out/f_configs.err
