#!/bin/bash
# round 2, call H: packed generator + relaxed epilogue waits + narrow Q^T F kernel
mkdir -p gpurun_out
timeout 300 build/umma_probe 50000 200 2048 i8 > gpurun_out/probe5.txt 2>&1; echo "probe exit $?" >> gpurun_out/probe5.txt
grep -E "m=50000|m=20000|accuracy|layout tests|full tests" gpurun_out/probe5.txt
timeout 900 python -m pytest tests/test_zz_lowrank.py tests/test_cuda_parity.py -m gpu -q --maxfail=10 -rfEs --tb=short > gpurun_out/pytest_h.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_h.txt
tail -5 gpurun_out/pytest_h.txt
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5_h.json 2> gpurun_out/bench_cfg5_h.err; echo "exit $?" >> gpurun_out/bench_cfg5_h.err
python -c "
import json; j=json.load(open('gpurun_out/bench_cfg5_h.json')); print('cfg5', j['value'], j['ms_per_step'], j['setup_ms'], j['setup'], j['e2e']['value'], j['stage_ms'])"
export CPD_B200_NO_GRAPH=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gi_gram_kernel -s 2 -c 1 -f -o gpurun_out/prof_gi4 python tools/lr_prof.py lowrank 50000 > gpurun_out/prof_gi4.log 2>&1
