#!/usr/bin/env python
"""Low-rank non-rigid CPD on the device (BASELINE.json config 5: M = N = 50k, K = 200): set-up time (range finder) and time
per EM iteration, with the dense path beside it where it still fits.   usage: lowrank_timing.py [K] [M ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair

rank = int(sys.argv[1]) if len(sys.argv) > 1 else 200
sizes = [int(a) for a in sys.argv[2:]] or [10000, 50000]
f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
for n in sizes:
    src, _ = synthetic_pair(n)
    tgt = np.ascontiguousarray(src + 0.03 * np.sin(2 * np.pi * src.dot(f)))
    h = _cabi.Handle(3)
    h.set_source(src); h.set_target(tgt)
    s2 = h.sigma2_init()
    for piters in (2, 0):
        h.sync(); t0 = time.perf_counter(); h.nonrigid_lowrank_begin(2.0, 2.0, s2, 0.0, rank, piters, 0); h.sync(); t_b = time.perf_counter() - t0
        l0 = h.launch_count()
        sig = [h.nonrigid_step()]
        t0 = time.perf_counter()
        for _ in range(5):
            sig.append(h.nonrigid_step())
        dt = (time.perf_counter() - t0) / 5
        print("M=N=%6d K=%d power_iters=%d: set-up %.1f ms, %.2f ms/iteration (%d launches/iteration), sigma2 %s" % (
            n, rank, piters, t_b * 1e3, dt * 1e3, (h.launch_count() - l0) // 6, ["%.5g" % s for s in sig[::2]]), flush=True)
    if n <= 20000:
        t0 = time.perf_counter(); h.nonrigid_begin(2.0, 2.0, s2, 0.0); h.sync(); t_g = time.perf_counter() - t0
        sig = [h.nonrigid_step()]
        t0 = time.perf_counter()
        for _ in range(2):
            sig.append(h.nonrigid_step())
        print("            dense: G build %.1f ms, %.1f ms/iteration, sigma2 %s" % (t_g * 1e3, (time.perf_counter() - t0) / 2 * 1e3, ["%.5g" % s for s in sig]), flush=True)
    h.close()
