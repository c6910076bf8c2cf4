#!/usr/bin/env python
"""the reference's own CPU path, timed directly at N = M in {1k, 4k, 10k} (one EM iteration = transform +
expectation_step + maximization_step of the UNMODIFIED probreg/cpd.py, loaded as tests/golden/make_golden.py loads it), next to
the oracle port that bench.py uses as `cpu_baseline`.  Needs /root/reference, so it runs in the build container only; the
output is kept under profiles/.   usage: python tools/cpu_reference_timing.py [sizes...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as mg  # noqa: E402
from oracle import cpd_oracle as orc  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [1000, 4000, 10000]
ref_cpd, _ = mg.load_reference()
try:
    import threadpoolctl
    blas = threadpoolctl.threadpool_info()
    threads = max([b.get("num_threads", 1) for b in blas] or [1])
except Exception:
    threads = "?"
print("host: %d cores (os.cpu_count), BLAS threads %s, numpy %s" % (os.cpu_count(), threads, np.__version__))
for n in sizes:
    src, tgt = orc.synthetic_pair(n)
    reg = ref_cpd.RigidCPD(src)
    res = reg._initialize(tgt)
    t_e = t_m = 0.0
    iters = 3 if n <= 4000 else 2
    for _ in range(iters):
        ts = res.transformation.transform(src)
        t0 = time.perf_counter()
        es = reg.expectation_step(ts, tgt, res.sigma2, 0.0)
        t1 = time.perf_counter()
        res = reg.maximization_step(tgt, es, res.sigma2)
        t2 = time.perf_counter()
        t_e += t1 - t0
        t_m += t2 - t1
    t_e /= iters
    t_m /= iters
    t0 = time.perf_counter()
    orc.expectation_step(ts, tgt, res.sigma2, 0.0)
    t_o = time.perf_counter() - t0
    print("N = M = %6d: reference E-step %.3f s (%.1f ns/pair), M-step %.4f s -> %.3f it/s;  oracle port E-step %.3f s (%.1f ns/pair)" % (
        n, t_e, t_e / (n * n) * 1e9, t_m, 1.0 / (t_e + t_m), t_o, t_o / (n * n) * 1e9), flush=True)
