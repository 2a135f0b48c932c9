#!/bin/bash
# GPU session O (re-entry check): parity tests, smoke, default bench, reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/o_smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/o_pytest.log
tail -6 gpurun_out/o_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/o_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/o_smoke.log
tail -3 gpurun_out/o_smoke.log
( time timeout 600 python bench.py ) > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err
cut -c1-1500 gpurun_out/o_bench.json; tail -4 gpurun_out/o_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/o_bench_ref.json 2> gpurun_out/o_bench_ref.err
cut -c1-800 gpurun_out/o_bench_ref.json; tail -4 gpurun_out/o_bench_ref.err
