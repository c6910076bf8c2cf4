#!/bin/bash
# ncu evidence: launch list of a short bench run + one full capture of the two E-step kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass[12]_kernel -s 2 -c 2 -f \
    -o gpurun_out/prof python tools/prof_step.py 100000 2 > gpurun_out/prof.log 2>&1
tail -3 gpurun_out/prof.log; wc -l gpurun_out/launches.csv
