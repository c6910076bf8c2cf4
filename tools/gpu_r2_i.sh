#!/bin/bash
# round 2, call I: converged MMA-issue warp (uniform datapath), 256-record E-step stages
mkdir -p gpurun_out
timeout 300 build/umma_probe 50000 200 2048 > gpurun_out/probe6.txt 2>&1; echo "probe exit $?" >> gpurun_out/probe6.txt
grep -E "m=50000|m=20000|layout tests|full tests" gpurun_out/probe6.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -rfEs --tb=short > gpurun_out/pytest_i.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_i.txt
tail -6 gpurun_out/pytest_i.txt
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5_i.json 2> gpurun_out/bench_cfg5_i.err; echo "exit $?" >> gpurun_out/bench_cfg5_i.err
python -c "
import json; j=json.load(open('gpurun_out/bench_cfg5_i.json')); print('cfg5', j['value'], j['ms_per_step'], j['setup_ms'], j['setup'], j['e2e']['value'], j['stage_ms'])"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-also > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "exit $?" >> gpurun_out/bench_i.err
python -c "
import json; j=json.load(open('gpurun_out/bench_i.json')); print('cfg2', j['value'], j['ms_per_step'], j['e2e']['value'], j['stage_ms'], j['roofline']['frac'])"
timeout 300 python tools/shard_shape.py > gpurun_out/shard_shapes.txt 2>&1; cat gpurun_out/shard_shapes.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_i.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_i.txt; tail -5 gpurun_out/smoke_i.txt
export CPD_B200_NO_GRAPH=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gi_gram_kernel -s 2 -c 1 -f -o gpurun_out/prof_gi5 python tools/lr_prof.py lowrank 50000 > gpurun_out/prof_gi5.log 2>&1
