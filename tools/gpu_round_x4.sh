#!/bin/bash
# 4-GPU check of the final code: bench.py and the config-3 ensemble at 1 / 2 / 4 ranks.
mkdir -p gpurun_out
rm -f gpurun_out/x4_bench.jsonl gpurun_out/x4_ensemble.jsonl
run_n() {
  n=$1; shift
  if [ "$n" = 1 ]; then timeout 600 python "$@"; else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; fi
}
for n in 1 2 4; do
  run_n $n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline 2>>gpurun_out/x4.err | tail -1 >> gpurun_out/x4_bench.jsonl
  run_n $n tools/bench_ensemble.py --B 1024 --steps 5 2>>gpurun_out/x4.err | tail -1 >> gpurun_out/x4_ensemble.jsonl
done
python - <<'PY'
import json
for f in ("gpurun_out/x4_bench.jsonl", "gpurun_out/x4_ensemble.jsonl"):
    for l in open(f):
        try: d = json.loads(l)
        except Exception: print("bad line:", l[:200]); continue
        print(f.split("/")[-1], d.get("n_gpus"), d.get("value"), d.get("unit"), (d.get("e2e") or {}).get("value"))
PY
