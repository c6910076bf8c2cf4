#!/bin/bash
# round 2, call A: the GPU suite with xfail marks ignored (--runxfail) and long tracebacks; low-rank timing.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --runxfail -rfEs --tb=long > gpurun_out/pytest_runxfail.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_runxfail.txt
timeout 300 python tools/lowrank_timing.py 200 50000 > gpurun_out/lowrank_timing.txt 2>&1
tail -40 gpurun_out/pytest_runxfail.txt; cat gpurun_out/lowrank_timing.txt
