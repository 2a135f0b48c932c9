#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/y_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/y_pytest.log
tail -30 gpurun_out/y_pytest.log
