#!/bin/bash
# round 2, call F: int8 product with staged j-points, new panel kernels; tests; launch list + ncu of the product
mkdir -p gpurun_out
timeout 300 build/umma_probe 50000 200 2048 i8 > gpurun_out/i8_probe3.txt 2>&1; echo "probe exit $?" >> gpurun_out/i8_probe3.txt
tail -6 gpurun_out/i8_probe3.txt
timeout 900 python -m pytest tests/test_zz_lowrank.py -m gpu -q --maxfail=10 -rfEs --tb=short > gpurun_out/pytest_f.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_f.txt
tail -8 gpurun_out/pytest_f.txt
python - > gpurun_out/lowrank_setup3.txt 2>&1 <<'PY'
import sys, time, os
sys.path.insert(0, '.')
import numpy as np
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair
src, _ = synthetic_pair(50000)
f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
tgt = np.ascontiguousarray(src + 0.03 * np.sin(2 * np.pi * src.dot(f)))
h = _cabi.Handle(3); h.set_source(src); h.set_target(tgt); s2 = h.sigma2_init()
h.set_profiling(True)
for rep in range(3):
    h.sync(); t0 = time.perf_counter(); h.nonrigid_lowrank_begin(2.0, 2.0, s2, 0.0, 200, 2, 0); h.sync(); dt = time.perf_counter() - t0
    print("set-up %.1f ms wall; phases:" % (dt * 1e3), h.lowrank_setup_times(), "sigma2_1 %.9g" % h.nonrigid_step(), flush=True)
for _ in range(3): h.nonrigid_step()
print("stage times of an iteration:", h.stage_times())
PY
cat gpurun_out/lowrank_setup3.txt
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5_f.json 2> gpurun_out/bench_cfg5_f.err; echo "exit $?" >> gpurun_out/bench_cfg5_f.err
head -c 1200 gpurun_out/bench_cfg5_f.json; tail -2 gpurun_out/bench_cfg5_f.err
export CPD_B200_NO_GRAPH=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lr_launches2.csv python tools/lr_prof.py lowrank 50000 > gpurun_out/lr_under_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gi_gram_kernel -s 2 -c 1 -f -o gpurun_out/prof_gi2 python tools/lr_prof.py lowrank 50000 > gpurun_out/prof_gi2.log 2>&1
python tools/launch_shares.py gpurun_out/lr_launches2.csv | head -16
