#!/bin/bash
# What the driver runs at round end (GPU tests, smoke(), bench.py, reference arm), for one gpurun call:
#   gpurun --timeout 2400 -- 'bash tools/gpu_check.sh [tag]'
# Outputs land in gpurun_out/<tag>_*.
tag=${1:-check}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
tail -3 gpurun_out/${tag}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-3000 gpurun_out/${tag}_bench.json; tail -4 gpurun_out/${tag}_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
cut -c1-400 gpurun_out/${tag}_bench_ref.json; tail -4 gpurun_out/${tag}_bench_ref.err
