#!/bin/bash
# 2-GPU session: trajectory-sharded scaling of the headline bench and of the config-3 ensemble
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/l_gpus.txt
python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/l_bench_1.json 2> gpurun_out/l_err.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/l_bench_2.json 2>> gpurun_out/l_err.txt
python tools/bench_ensemble.py --B 1024 > gpurun_out/l_ens_1.json 2>> gpurun_out/l_err.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 tools/bench_ensemble.py --B 1024 > gpurun_out/l_ens_2.json 2>> gpurun_out/l_err.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "ensemble or nccl or two" > gpurun_out/l_pytest.log 2>&1; tail -2 gpurun_out/l_pytest.log
for f in gpurun_out/l_bench_1.json gpurun_out/l_bench_2.json gpurun_out/l_ens_1.json gpurun_out/l_ens_2.json; do grep -o '"value": [0-9.]*' $f | head -1; done
tail -3 gpurun_out/l_err.txt
