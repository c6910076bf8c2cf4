#!/bin/bash
# low-rank path on one GPU: its tests, the config-5 bench line, and the launch list of two iterations (kernel shares)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_lowrank.py -m gpu -q --maxfail=10 -rfEs --tb=short > gpurun_out/pytest_lr.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_lr.txt
tail -3 gpurun_out/pytest_lr.txt
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "exit $?" >> gpurun_out/bench_cfg5.err
python -c "
import json; j=json.load(open('gpurun_out/bench_cfg5.json')); print('cfg5', j['value'], j['ms_per_step'], j['setup_ms'], j['setup_cold_ms'], j['setup'], j['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_lr.csv python tools/lr_prof.py lowrank > gpurun_out/lr_prof.log 2>&1
python tools/launch_shares.py gpurun_out/launches_lr.csv > gpurun_out/launch_shares_lr.txt 2>&1; grep -E "lr_inner_sym|lr_spd|lr_pchol|lr_rotate|lr_llt|lr_merge_sym|getrf|lr_apply|lr_inner_narrow|pass1|pass2|^total" gpurun_out/launch_shares_lr.txt
