mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_interfaces.py tests/test_gpu_ensemble.py -m gpu -x -q ) > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_pytest.log
tail -8 gpurun_out/r2m_pytest.log | cut -c1-300
