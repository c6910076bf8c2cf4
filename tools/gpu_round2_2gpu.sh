#!/bin/bash
# round 2, 2-GPU call: NCCL + peer-mailbox parity at world 2, the bench line at N=2 (graph + staged M-step + one-sync sigma2 init)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/nvsmi_2gpu.txt
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -rfEs --tb=short > gpurun_out/pytest_2gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_2gpu.txt
tail -6 gpurun_out/pytest_2gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "exit $?" >> gpurun_out/bench_2gpu.err
python -c "
import json; j=json.load(open('gpurun_out/bench_2gpu.json')); print('N=2', j['value'], j['ms_per_step'], j['e2e']['value'], j['stage_ms'], j['config'].get('sharded_equals_single_gpu'), j['result_check'])"
tail -3 gpurun_out/bench_2gpu.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-also > gpurun_out/bench_1of2.json 2> gpurun_out/bench_1of2.err
python -c "
import json; j=json.load(open('gpurun_out/bench_1of2.json')); print('N=1', j['value'], j['ms_per_step'], j['e2e']['value'], j['stage_ms'])"
