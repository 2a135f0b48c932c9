#!/bin/bash
# GPU session V: full parity suite + all secondary configs + default bench after the SpMM change.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/v_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/v_pytest.log
tail -15 gpurun_out/v_pytest.log
timeout 900 python tools/bench_configs.py --configs 1,3,5 > gpurun_out/v_configs.jsonl 2> gpurun_out/v_configs.err
cut -c1-420 gpurun_out/v_configs.jsonl; tail -3 gpurun_out/v_configs.err
timeout 600 python bench.py > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; cut -c1-300 gpurun_out/v_bench.json
