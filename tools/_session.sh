mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiled" ) > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
tail -3 gpurun_out/r2i_pytest.log
timeout 300 python tools/bench_ensemble.py --B 1024 --steps 5 --warmup 2 2>gpurun_out/r2i_ens.err | tail -1 > gpurun_out/r2i_ens.json; cat gpurun_out/r2i_ens.json; tail -3 gpurun_out/r2i_ens.err
