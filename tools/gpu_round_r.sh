#!/bin/bash
# GPU session R: full parity suite incl. the interface checks on device types.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r_pytest.log
tail -40 gpurun_out/r_pytest.log
