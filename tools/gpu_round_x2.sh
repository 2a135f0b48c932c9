#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/x2.err | tail -1 > gpurun_out/x2_bench.json
python -c "
import json; d = json.load(open('gpurun_out/x2_bench.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['scaling'], d.get('cpu_baseline') is not None)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>>gpurun_out/x2.err | tail -1 | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_ensemble.py -m gpu -q 2>&1 | tail -2
tail -3 gpurun_out/x2.err
