#!/bin/bash
# bench.py (config 2 replicas + config-3 ensemble object) at 1 / 2 / ... ranks of one box:
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_scale.sh 1 2 4 8'
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/scale_gpus.txt
for n in "$@"; do
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 --no-cpu-baseline 2>>gpurun_out/scale.err | tail -1 >> gpurun_out/scale_bench.jsonl
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
       bench.py --gpus $n --no-cpu-baseline 2>>gpurun_out/scale.err | tail -1 >> gpurun_out/scale_bench.jsonl; fi
done
python - <<'PY'
import json
for l in open("gpurun_out/scale_bench.jsonl"):
    try: d = json.loads(l)
    except Exception: print("bad line:", l[:200]); continue
    ens = d.get("ensemble") or {}
    print(d.get("n_gpus"), d.get("value"), d.get("unit"), (d.get("e2e") or {}).get("value"), "| ensemble", ens.get("value"), ens.get("unit"), ens.get("roofline_frac_per_gpu"))
PY
tail -5 gpurun_out/scale.err
