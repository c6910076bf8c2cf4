#!/bin/bash
# round 2, call J: dead-warp skip in pass 1/2 (512-record stages again): N=1 bench, shard shapes, parity
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-also > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "exit $?" >> gpurun_out/bench_j.err
python -c "
import json; j=json.load(open('gpurun_out/bench_j.json')); print('cfg2', j['value'], j['ms_per_step'], j['e2e']['value'], j['stage_ms'], j['roofline']['frac'])"
timeout 300 python tools/shard_shape.py > gpurun_out/shard_shapes_j.txt 2>&1; cat gpurun_out/shard_shapes_j.txt
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_cuda_edges.py tests/test_zz_baseline_configs.py -m gpu -q --maxfail=10 -rfEs --tb=short > gpurun_out/pytest_j.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_j.txt
tail -4 gpurun_out/pytest_j.txt
