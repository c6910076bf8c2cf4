#!/bin/bash
# One GPU visit: parity tests, smoke, bench.  Usage (via gpurun): bash tools/gpu_round.sh [pytest-args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 "$@" > gpurun_out/pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/pytest.txt; cat gpurun_out/smoke.txt | tail -3; cat gpurun_out/bench.json | head -c 3000; tail -3 gpurun_out/bench.err
