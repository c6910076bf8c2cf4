#!/bin/bash
# round 2, call E: reworked int8 product (bulk-copied stage images), low-rank tests, ncu evidence of the low-rank / BCPD / shard kernels
mkdir -p gpurun_out
timeout 300 build/umma_probe 50000 200 2048 i8 > gpurun_out/i8_probe2.txt 2>&1; echo "probe exit $?" >> gpurun_out/i8_probe2.txt
cat gpurun_out/i8_probe2.txt
timeout 900 python -m pytest tests/test_zz_lowrank.py tests/test_cuda_parity.py -m gpu -q --maxfail=10 -rfEs --tb=short > gpurun_out/pytest_e.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_e.txt
tail -15 gpurun_out/pytest_e.txt
python - > gpurun_out/lowrank_setup2.txt 2>&1 <<'PY'
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
PY
cat gpurun_out/lowrank_setup2.txt
export CPD_B200_NO_GRAPH=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lr_launches.csv python tools/lr_prof.py lowrank 50000 > gpurun_out/lr_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gi_gram_kernel -c 1 -f -o gpurun_out/prof_gi python tools/lr_prof.py lowrank 50000 > gpurun_out/prof_gi.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lr_inner_kernel\|lr_panel_update_kernel\|lr_panel_apply_kernel -s 40 -c 6 -f -o gpurun_out/prof_lr python tools/lr_prof.py lowrank 50000 > gpurun_out/prof_lr.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass[12]_kernel -s 2 -c 2 -f -o gpurun_out/prof_wgt python tools/lr_prof.py bcpd 100000 > gpurun_out/prof_wgt.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass[12]_kernel\|moments_kernel -s 4 -c 3 -f -o gpurun_out/prof_shard python tools/lr_prof.py shard 100000 8 > gpurun_out/prof_shard.log 2>&1
ls -la gpurun_out/*.ncu-rep; wc -l gpurun_out/lr_launches.csv
