#!/bin/bash
# round 2, the 8-GPU call: multi-rank parity (2/4/8 ranks), BASELINE config 2 at N = 8 and 4, BASELINE config 4 (rigid 1M, 8 ranks)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/nvsmi_8gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "exit $?" >> gpurun_out/bench_8gpu.err
python -c "
import json; j=json.load(open('gpurun_out/bench_8gpu.json')); print('N=8 cfg2', j['value'], j['ms_per_step'], j['e2e']['value'], j['stage_ms'], j['config'].get('sharded_equals_single_gpu'))"
tail -2 gpurun_out/bench_8gpu.err
timeout 1200 $TR --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 --config 4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4_8gpu.json 2> gpurun_out/bench_cfg4_8gpu.err; echo "exit $?" >> gpurun_out/bench_cfg4_8gpu.err
python -c "
import json; j=json.load(open('gpurun_out/bench_cfg4_8gpu.json')); print('N=8 cfg4', j['value'], j['ms_per_step'], j['e2e']['value'], j['stage_ms'], j['roofline']['frac'], j['config'].get('sharded_equals_single_gpu'))"
tail -2 gpurun_out/bench_cfg4_8gpu.err
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 $TR --nproc-per-node 4 --master-port 29703 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "exit $?" >> gpurun_out/bench_4gpu.err
python -c "
import json; j=json.load(open('gpurun_out/bench_4gpu.json')); print('N=4 cfg2', j['value'], j['ms_per_step'], j['e2e']['value'], j['config'].get('sharded_equals_single_gpu'))"
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -rfEs --tb=short > gpurun_out/pytest_8gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_8gpu.txt
tail -5 gpurun_out/pytest_8gpu.txt
