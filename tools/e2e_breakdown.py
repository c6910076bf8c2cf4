#!/usr/bin/env python
"""Where the per-call overhead of registration(maxiter=1) goes: host-timed phases (each followed by a stream sync, so they do not
overlap) for the full 100k x 100k problem and for the shard a rank holds in an 8-rank run (100k sources x 12.5k targets)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair

def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True); t.numpy()[...] = a; return t.numpy()

n = 100000
src, tgt = synthetic_pair(n)
src, tgt = pinned(src), pinned(tgt)
for world in (1, 8):
    sh = tgt[: n // world]
    h = _cabi.Handle(3)
    origin = tgt[:: max(1, n // 1024)].mean(axis=0)
    acc = {}
    def phase(name, fn):
        h.sync(); t0 = time.perf_counter(); r = fn(); h.sync(); acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0); return r
    for rep in range(12):
        if rep == 2: acc.clear()
        phase("set_source", lambda: h.set_source(src))
        phase("set_target", lambda: h.set_target(sh, n_global=n, frame_origin=origin))
        s2 = phase("sigma2_init", lambda: h.sigma2_init())
        phase("set_state", lambda: h.set_state(_cabi.TF_RIGID, True, 0.0, np.identity(3), np.zeros(3), 1.0, s2, 1.0 + n * 1.5 * np.log(s2)))
        phase("em_step+read", lambda: h.em_step())
    tot = sum(acc.values()) / 10
    print("world %d: total %.3f ms per call (phases synchronised): " % (world, tot * 1e3) + ", ".join("%s %.0f us" % (k, v / 10 * 1e6) for k, v in acc.items()))
    t0 = time.perf_counter()
    for rep in range(10):
        h.set_source(src); h.set_target(sh, n_global=n, frame_origin=origin); s2 = h.sigma2_init()
        h.set_state(_cabi.TF_RIGID, True, 0.0, np.identity(3), np.zeros(3), 1.0, s2, 1.0); h.em_step()
    print("          unsynchronised sequence: %.3f ms per call" % ((time.perf_counter() - t0) / 10 * 1e3))
    h.close()
