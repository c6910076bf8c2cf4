#!/bin/bash
# the round-end sequence on one GPU (tests, smoke, reference arm, bench line): profiles/r2_pytest_gpu_final_1gpu.txt, r2_smoke.txt,
# r2_bench_reference_arm.json, r2_bench_1gpu_final.json
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -rfEs --tb=short > gpurun_out/pytest_final.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_final.txt
tail -6 gpurun_out/pytest_final.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke_final.txt; tail -4 gpurun_out/smoke_final.txt
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?" >> gpurun_out/bench_final.err
python -c "
import json; j=json.load(open('gpurun_out/bench_final.json')); print('cfg2', j['value'], j['ms_per_step'], j['e2e']['value'], j['stage_ms'], j['roofline']['frac']); print(json.dumps(j['config']['also'])[:1500]); print(j['cpu_baseline']['value'], j['cpu_baseline']['kind'])
r=json.load(open('gpurun_out/bench_ref_final.json')); print('ref', r['value'], r['cpu_baseline']['kind'], r['config'])"
tail -3 gpurun_out/bench_final.err
