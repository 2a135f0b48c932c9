mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bundle or newton or arnoldi" ) > gpurun_out/r2l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2l_pytest.log
tail -30 gpurun_out/r2l_pytest.log
