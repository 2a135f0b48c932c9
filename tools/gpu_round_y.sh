#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge_shapes" ) > gpurun_out/y_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/y_pytest.log
tail -40 gpurun_out/y_pytest.log
