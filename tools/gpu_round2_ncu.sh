#!/bin/bash
# round-2 ncu evidence for the bench command: launch list of a short bench run (graph launches: kernel nodes are profiled individually)
# and a full capture of the two E-step kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_shares.py gpurun_out/launches_r2.csv > gpurun_out/launch_shares_r2.txt; head -8 gpurun_out/launch_shares_r2.txt
CPD_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass[12]_kernel -s 2 -c 2 -f \
    -o gpurun_out/prof_estep_r2 python tools/prof_step.py 100000 2 > gpurun_out/prof_estep_r2.log 2>&1
ls -la gpurun_out/prof_estep_r2.ncu-rep
