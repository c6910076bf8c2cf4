#!/usr/bin/env python
"""Workloads for ncu (one per invocation):  lr_prof.py lowrank [points] | bcpd [points] | shard [points] [world]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair

what = sys.argv[1] if len(sys.argv) > 1 else "lowrank"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
src, tgt = synthetic_pair(n)
if what == "lowrank":
    f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
    tgt = np.ascontiguousarray(src + 0.03 * np.sin(2 * np.pi * src.dot(f)))
    h = _cabi.Handle(3); h.set_source(src); h.set_target(tgt)
    h.nonrigid_lowrank_begin(2.0, 2.0, h.sigma2_init(), 0.0, 200, 2, 0)
    print([h.nonrigid_step() for _ in range(2)])
elif what == "bcpd":
    rng = np.random.default_rng(1)
    h = _cabi.Handle(3); h.set_source(src); h.set_target(tgt)
    alpha, sdiag = rng.dirichlet(np.ones(n)), rng.uniform(0.0, 1e-3, n)
    for _ in range(2):
        out = h.bcpd_estep(src, 1.0, alpha, sdiag, 2e-3, 0.1)
    print(out[3])
elif what == "shard":
    world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    h = _cabi.Handle(3); h.set_source(src)
    h.set_target(tgt[: n // world], n_global=n, frame_origin=tgt.mean(0))
    h.set_state(_cabi.TF_RIGID, True, 0.0, np.identity(3), np.zeros(3), 1.0, 0.11, 1.0)
    os.environ.setdefault("CPD_B200_NO_GRAPH", "1")
    for _ in range(4):
        out = h.em_step()
    print(out[3])
