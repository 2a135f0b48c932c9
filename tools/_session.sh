mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiled" ) > gpurun_out/r2t_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2t_pytest.log
tail -25 gpurun_out/r2t_pytest.log | cut -c1-250
