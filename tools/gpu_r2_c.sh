#!/bin/bash
# round 2, call C: full GPU suite (no xfail marks left), smoke, the bench line with config.also, low-rank set-up phases
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -rfEs --tb=short > gpurun_out/pytest_c.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_c.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt
timeout 300 build/umma_probe 50000 200 1024 > gpurun_out/umma_probe2.txt 2>&1; echo "probe exit $?" >> gpurun_out/umma_probe2.txt
python - > gpurun_out/lowrank_setup.txt 2>&1 <<'PY'
import sys, time, os
sys.path.insert(0, '.')
import numpy as np
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair
src, _ = synthetic_pair(50000)
f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
tgt = np.ascontiguousarray(src + 0.03 * np.sin(2 * np.pi * src.dot(f)))
for mode in ("default", "columnwise", "simt"):
    os.environ.pop("CPD_B200_LR_ORTH", None)
    if mode == "columnwise": os.environ["CPD_B200_LR_ORTH"] = "columnwise"
    h = _cabi.Handle(3); h.set_source(src); h.set_target(tgt); s2 = h.sigma2_init()
    h.set_profiling(True)
    for rep in range(2):
        h.sync(); t0 = time.perf_counter(); h.nonrigid_lowrank_begin(2.0, 2.0, s2, 0.0, 200, 2, 0); h.sync(); dt = time.perf_counter() - t0
        print(mode, "set-up %.1f ms wall; phases (products, orthonormalisations, core) ms:" % (dt * 1e3), [round(x, 3) for x in h.lowrank_setup_times()], "sigma2_1 %.9g" % h.nonrigid_step(), flush=True)
    h.close()
PY
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench exit $?" >> gpurun_out/bench_c.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c.json 2> gpurun_out/bench_ref_c.err
tail -30 gpurun_out/pytest_c.txt; tail -3 gpurun_out/smoke.txt; cat gpurun_out/umma_probe2.txt | tail -12; cat gpurun_out/lowrank_setup.txt; head -c 3000 gpurun_out/bench_c.json; tail -5 gpurun_out/bench_c.err; head -c 1500 gpurun_out/bench_ref_c.json
