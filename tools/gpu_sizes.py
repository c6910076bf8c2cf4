#!/usr/bin/env python
"""BASELINE.json's larger configurations, a few iterations each: affine 250k and rigid 1M on one GPU (properties only:
the oracle cannot run these sizes) -- finite results, sum(p1) == N, monotone sigma2, timing per iteration."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair

for name, n, kind, tf in [("affine 250k", 250000, "affine", _cabi.TF_AFFINE), ("rigid 1M", 1000000, "rigid", _cabi.TF_RIGID)]:
    src, tgt = synthetic_pair(n, kind)
    h = _cabi.Handle(3)
    t0 = time.perf_counter(); h.set_source(src); h.set_target(tgt); s2 = h.sigma2_init(); t_up = time.perf_counter() - t0
    h.set_state(tf, True, 0.0, np.identity(3), np.zeros(3), 1.0, s2, 1.0 + n * 1.5 * np.log(s2))
    h.em_step()
    sig = [s2]
    h.timer_start()
    for _ in range(4):
        out = h.em_step(); sig.append(out[3])
    ms = h.timer_stop() / 4
    assert np.all(np.isfinite(out[0])) and all(b < a for a, b in zip(sig[1:], sig[2:])), sig
    assert abs(out[5] - n) < 1e-6 * n, out[5]
    print("%-12s upload+init %.1f ms, %.2f ms/iteration (%.3f it/s), sigma2 %s" % (name, t_up * 1e3, ms, 1e3 / ms, ["%.4g" % s for s in sig]), flush=True)
    h.close()
