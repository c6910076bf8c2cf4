#!/bin/bash
# First GPU visit of round 2 (one gpurun call, ~12-15 min of box time): everything written without a GPU in round 1 gets its
# first hardware run, plus the numbers NEXT.md asks for.   usage:  gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
# 1. the verified tests first, then the pending ones (file names sort them last); -rxX lists every XPASS / XFAIL with its reason
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -rxXfE > gpurun_out/pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.txt
# 2. the bench line (finalize 1 gained a parameter; the hot kernels' SASS is unchanged)
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
# 3. BASELINE configuration 5 and the dense path beside it
timeout 600 python tools/lowrank_timing.py 200 10000 50000 > gpurun_out/lowrank_timing.txt 2>&1
# 4. ncu: per-launch times of one low-rank set-up + 2 iterations at 20k / K = 200, then a full capture of the two kernels that matter
cat > /tmp/lr_prof.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair
n = int(sys.argv[1])
src, _ = synthetic_pair(n)
f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
tgt = np.ascontiguousarray(src + 0.03 * np.sin(2 * np.pi * src.dot(f)))
h = _cabi.Handle(3); h.set_source(src); h.set_target(tgt)
h.nonrigid_lowrank_begin(2.0, 2.0, h.sigma2_init(), 0.0, 200, 1, 0)
print([h.nonrigid_step() for _ in range(2)])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lr_launches.csv \
    python /tmp/lr_prof.py 20000 > gpurun_out/lr_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lr_gram_apply_kernel\|lr_inner_kernel -c 3 -f \
    -o gpurun_out/lr_prof python /tmp/lr_prof.py 20000 > gpurun_out/lr_prof.log 2>&1
tail -25 gpurun_out/pytest.txt; tail -2 gpurun_out/smoke.txt; head -c 1500 gpurun_out/bench.json; echo; cat gpurun_out/lowrank_timing.txt; wc -l gpurun_out/lr_launches.csv
