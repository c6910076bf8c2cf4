#!/bin/bash
# round 2, call G: shared-memory address space fix (LDS/STS), staged M-step; full suite + timings
mkdir -p gpurun_out
timeout 300 build/umma_probe 50000 200 2048 > gpurun_out/probe4.txt 2>&1; echo "probe exit $?" >> gpurun_out/probe4.txt
grep -E "m=50000|layout tests|full tests" gpurun_out/probe4.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -rfEs --tb=short > gpurun_out/pytest_g.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_g.txt
tail -8 gpurun_out/pytest_g.txt
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5_g.json 2> gpurun_out/bench_cfg5_g.err; echo "exit $?" >> gpurun_out/bench_cfg5_g.err
python -c "
import json; j=json.load(open('gpurun_out/bench_cfg5_g.json')); print('cfg5', j['value'], j['ms_per_step'], j['setup_ms'], j['setup'], j['e2e']['value'], j['stage_ms'])"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-also > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "exit $?" >> gpurun_out/bench_g.err
python -c "
import json; j=json.load(open('gpurun_out/bench_g.json')); print('cfg2', j['value'], j['ms_per_step'], j['e2e']['value'], j['stage_ms'], j['roofline']['frac'])"
export CPD_B200_NO_GRAPH=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gi_gram_kernel -s 2 -c 1 -f -o gpurun_out/prof_gi3 python tools/lr_prof.py lowrank 50000 > gpurun_out/prof_gi3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lr_launches3.csv python tools/lr_prof.py lowrank 50000 > gpurun_out/lr_under_ncu3.log 2>&1
python tools/launch_shares.py gpurun_out/lr_launches3.csv > gpurun_out/lr_shares3.txt; head -14 gpurun_out/lr_shares3.txt
