#!/bin/bash
# round 2, call D: int8 exact product bring-up, then the GPU suite on the library that now uses it (+ CUDA-graph EM step, 2 ranks on 1 GPU)
mkdir -p gpurun_out
timeout 300 build/umma_probe 50000 200 2048 i8 > gpurun_out/i8_probe.txt 2>&1; echo "probe exit $?" >> gpurun_out/i8_probe.txt
cat gpurun_out/i8_probe.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -rfEs --tb=short > gpurun_out/pytest_d.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_d.txt
tail -40 gpurun_out/pytest_d.txt
python - > gpurun_out/lowrank_setup.txt 2>&1 <<'PY'
import sys, time, os
sys.path.insert(0, '.')
import numpy as np
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair
src, _ = synthetic_pair(50000)
f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
tgt = np.ascontiguousarray(src + 0.03 * np.sin(2 * np.pi * src.dot(f)))
for mode in ("default", "columnwise"):
    os.environ.pop("CPD_B200_LR_ORTH", None)
    if mode == "columnwise": os.environ["CPD_B200_LR_ORTH"] = "columnwise"
    h = _cabi.Handle(3); h.set_source(src); h.set_target(tgt); s2 = h.sigma2_init()
    h.set_profiling(True)
    for rep in range(2):
        h.sync(); t0 = time.perf_counter(); h.nonrigid_lowrank_begin(2.0, 2.0, s2, 0.0, 200, 2, 0); h.sync(); dt = time.perf_counter() - t0
        print(mode, "set-up %.1f ms wall; phases:" % (dt * 1e3), h.lowrank_setup_times(), "sigma2_1 %.9g" % h.nonrigid_step(), flush=True)
    h.close()
PY
cat gpurun_out/lowrank_setup.txt
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "exit $?" >> gpurun_out/bench_cfg5.err
head -c 2500 gpurun_out/bench_cfg5.json; tail -3 gpurun_out/bench_cfg5.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-also > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "exit $?" >> gpurun_out/bench_d.err
head -c 1200 gpurun_out/bench_d.json; tail -3 gpurun_out/bench_d.err
