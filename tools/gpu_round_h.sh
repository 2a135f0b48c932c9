#!/bin/bash
mkdir -p gpurun_out
./tools/micro/dmma_peak > gpurun_out/h_dmma_peak.jsonl 2>&1; cat gpurun_out/h_dmma_peak.jsonl | tr '\n' ' '; echo
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/h_pytest.log
tail -3 gpurun_out/h_pytest.log
timeout 900 python tools/bench_configs.py --configs 3 --B 1024 >> gpurun_out/h_configs.jsonl 2>> gpurun_out/h_configs.err
python bench.py --no-cpu-baseline >> gpurun_out/h_configs.jsonl 2>> gpurun_out/h_configs.err
cut -c1-420 gpurun_out/h_configs.jsonl; tail -5 gpurun_out/h_configs.err
