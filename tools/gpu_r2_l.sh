#!/bin/bash
# round 2, call L: lr_inner with two points per LDS.128 (S product), ADVICE fixes; low-rank tests + config 5 + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_lowrank.py tests/test_cuda_parity.py -m gpu -q --maxfail=10 -rfEs --tb=short > gpurun_out/pytest_l.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_l.txt
tail -4 gpurun_out/pytest_l.txt
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5_l.json 2> gpurun_out/bench_cfg5_l.err; echo "exit $?" >> gpurun_out/bench_cfg5_l.err
python -c "
import json; j=json.load(open('gpurun_out/bench_cfg5_l.json')); print('cfg5', j['value'], j['ms_per_step'], j['setup_ms'], j['setup'], j['e2e']['value'], j['stage_ms'])"
export CPD_B200_NO_GRAPH=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lr_launches4.csv python tools/lr_prof.py lowrank 50000 > gpurun_out/lr_under_ncu4.log 2>&1
python tools/launch_shares.py gpurun_out/lr_launches4.csv > gpurun_out/lr_shares4.txt; head -12 gpurun_out/lr_shares4.txt
