#!/usr/bin/env python
"""Dense non-rigid CPD on the device: time per EM iteration at a few sizes (BASELINE.json config 5 is M = N = 50k)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair

sizes = [int(a) for a in sys.argv[1:]] or [5000, 10000, 20000]
f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
for n in sizes:
    src, _ = synthetic_pair(n)
    tgt = np.ascontiguousarray(src + 0.03 * np.sin(2 * np.pi * src.dot(f)))
    h = _cabi.Handle(3)
    h.set_source(src); h.set_target(tgt)
    s2 = h.sigma2_init()
    t0 = time.perf_counter(); h.nonrigid_begin(2.0, 2.0, s2, 0.0); h.sync(); t_g = time.perf_counter() - t0
    sig = [h.nonrigid_step()]
    t0 = time.perf_counter()
    for _ in range(3):
        sig.append(h.nonrigid_step())
    dt = (time.perf_counter() - t0) / 3
    print("M=N=%6d  G build %.1f ms, %.1f ms/iteration, sigma2 %s" % (n, t_g * 1e3, dt * 1e3, ["%.4g" % s for s in sig]), flush=True)
    h.close()
