mkdir -p gpurun_out
SEL="tiled or tfim_vs_oracle or operator_mul_selld or selld_uniform or expval or edge_shapes or bundle or config1_vs_oracle or leftright or check_normalization"
( time QPROP_SELL_KERNEL=ldg QPROP_TILE_SPIN_LIMIT=0 timeout 3000 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ensemble.py -m gpu -x -v -k "$SEL" ) > gpurun_out/r2_san_racecheck_ldg.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r2_san_racecheck_ldg.log
grep -E "FAILED|ERROR|hazard|SUMMARY|exit|Error|passed|failed" gpurun_out/r2_san_racecheck_ldg.log | tail -12 | cut -c1-200
grep -c PASSED gpurun_out/r2_san_racecheck_ldg.log
