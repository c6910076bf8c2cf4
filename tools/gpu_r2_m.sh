#!/bin/bash
# round 2, call M: A-operand-in-TMEM variant of the int8 product (probe: layout + full), library run with CPD_B200_LR_GRAM=i8ts, ncu of the S product
mkdir -p gpurun_out
timeout 300 build/umma_probe 50000 200 2048 i8 > gpurun_out/probe7.txt 2>&1; echo "probe exit $?" >> gpurun_out/probe7.txt
grep -E "m=50000|m=20000|m=3000|layout|full tests|timed out" gpurun_out/probe7.txt
CPD_B200_LR_GRAM=i8ts timeout 900 python -m pytest tests/test_zz_lowrank.py -m gpu -q --maxfail=10 -rfEs --tb=short > gpurun_out/pytest_m.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_m.txt
tail -4 gpurun_out/pytest_m.txt
CPD_B200_LR_GRAM=i8ts timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5_m.json 2> gpurun_out/bench_cfg5_m.err; echo "exit $?" >> gpurun_out/bench_cfg5_m.err
python -c "
import json; j=json.load(open('gpurun_out/bench_cfg5_m.json')); print('cfg5 i8ts', j['value'], j['ms_per_step'], j['setup_ms'], j['setup'], j['e2e']['value'])"
export CPD_B200_NO_GRAPH=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lr_inner_kernel -s 73 -c 1 -f -o gpurun_out/prof_s python tools/lr_prof.py lowrank 50000 > gpurun_out/prof_s.log 2>&1
CPD_B200_LR_GRAM=i8ts timeout 900 ncu --set full --clock-control none --import-source on -k regex:gi_gram_ts_kernel -s 2 -c 1 -f -o gpurun_out/prof_gits python tools/lr_prof.py lowrank 50000 > gpurun_out/prof_gits.log 2>&1
ls -la gpurun_out/prof_s.ncu-rep gpurun_out/prof_gits.ncu-rep
