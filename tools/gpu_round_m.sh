#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/m_pytest.log
tail -4 gpurun_out/m_pytest.log
timeout 600 python tools/bench_configs.py --configs 4 --liou-spins 10 --newton-steps 10 >> gpurun_out/m_configs.jsonl 2>> gpurun_out/m_configs.err
timeout 900 python tools/bench_configs.py --configs 4 --liou-spins 11 --newton-steps 5 >> gpurun_out/m_configs.jsonl 2>> gpurun_out/m_configs.err
cut -c1-400 gpurun_out/m_configs.jsonl; tail -5 gpurun_out/m_configs.err
