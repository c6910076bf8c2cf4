#!/bin/bash
# round 2, call N: TS int8 kernel with 16 generator warps
mkdir -p gpurun_out
timeout 300 build/umma_probe 50000 200 2048 i8 > gpurun_out/probe8.txt 2>&1; echo "probe exit $?" >> gpurun_out/probe8.txt
grep -E "m=50000|m=20000|m=3000|layout tests|full tests|timed out" gpurun_out/probe8.txt
CPD_B200_LR_GRAM=i8ts timeout 900 python -m pytest tests/test_zz_lowrank.py -m gpu -q --maxfail=10 -rfEs --tb=short > gpurun_out/pytest_n.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_n.txt
tail -3 gpurun_out/pytest_n.txt
export CPD_B200_NO_GRAPH=1
CPD_B200_LR_GRAM=i8ts timeout 900 ncu --set full --clock-control none --import-source on -k regex:gi_gram_ts_kernel -s 2 -c 1 -f -o gpurun_out/prof_gits2 python tools/lr_prof.py lowrank 50000 > gpurun_out/prof_gits2.log 2>&1
