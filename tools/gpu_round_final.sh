#!/bin/bash
# End-of-round check: what the driver runs (gpu tests, smoke, bench, reference arm) + secondary configs.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/f_pytest.log
tail -6 gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/f_smoke.log
tail -3 gpurun_out/f_smoke.log
( time timeout 600 python bench.py ) > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
cut -c1-2500 gpurun_out/f_bench.json; tail -4 gpurun_out/f_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
cut -c1-300 gpurun_out/f_bench_ref.json; tail -4 gpurun_out/f_bench_ref.err
timeout 900 python tools/bench_configs.py --configs 1,3,5 > gpurun_out/f_configs.jsonl 2> gpurun_out/f_configs.err
cut -c1-300 gpurun_out/f_configs.jsonl; tail -3 gpurun_out/f_configs.err
timeout 900 python tools/bench_configs.py --configs 4 --liou-spins 12 --newton-steps 3 > gpurun_out/f_config4_full.jsonl 2>> gpurun_out/f_configs.err
cut -c1-400 gpurun_out/f_config4_full.jsonl
