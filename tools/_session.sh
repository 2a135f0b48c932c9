mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ensemble.py -m gpu -x -q -k "tiled or ensemble" ) > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "config3 or config4 or config5" ) > gpurun_out/r2b_pytest2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest2.log
tail -15 gpurun_out/r2b_pytest2.log
( time timeout 600 python bench.py --no-cpu-baseline --no-sell-arm --min-seconds 0.3 ) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
    print(d['value'], json.dumps(d.get('ensemble'))[:1200])
except Exception as e: print('bench parse failed', e)
PY
tail -5 gpurun_out/r2b_bench.err
QPROP_NO_TILE=1 timeout 600 python bench.py --no-cpu-baseline --no-sell-arm --min-seconds 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NO_TILE', json.dumps(d.get('ensemble'))[:600])"
