#!/bin/bash
# 4-GPU call on the final build: NCCL + peer-mailbox parity at 2 and 4 ranks, the bench line at N = 4 and N = 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 4 --master-port 29711 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "exit $?" >> gpurun_out/bench_4gpu.err
python -c "
import json; j=json.load(open('gpurun_out/bench_4gpu.json')); print('N=4', j['value'], j['ms_per_step'], j['e2e']['value'], j['stage_ms'], j['config'].get('sharded_equals_single_gpu'))"
tail -2 gpurun_out/bench_4gpu.err
CUDA_VISIBLE_DEVICES=0,1 timeout 600 $TR --nproc-per-node 2 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "exit $?" >> gpurun_out/bench_2gpu.err
python -c "
import json; j=json.load(open('gpurun_out/bench_2gpu.json')); print('N=2', j['value'], j['ms_per_step'], j['e2e']['value'], j['stage_ms'], j['config'].get('sharded_equals_single_gpu'))"
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -rfEs --tb=short > gpurun_out/pytest_4gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_4gpu.txt
tail -5 gpurun_out/pytest_4gpu.txt
