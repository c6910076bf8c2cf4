#!/usr/bin/env python
"""Stage times on ONE GPU for the shard shape a rank sees in a W-rank run (N/W targets, all M sources)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair

n = 100000
src, tgt = synthetic_pair(n)
for world in (1, 2, 4, 8):
    h = _cabi.Handle(3)
    h.set_source(src)
    h.set_target(tgt[: n // world], n_global=n, frame_origin=tgt.mean(0))
    h.set_state(_cabi.TF_RIGID, True, 0.0, np.identity(3), np.zeros(3), 1.0, 0.11, 1.0)
    for _ in range(3):
        h.em_step(read=False)
    h.set_profiling(True)
    st = []
    for _ in range(8):
        h.em_step(read=False); st.append(h.stage_times())
    st = np.median(np.array(st), axis=0)
    print("shard 1/%d: pass1 %.3f pass2 %.3f fin1 %.3f fin2 %.3f mom %.3f total %.3f ms (ideal pass1 %.3f pass2 %.3f)" % (
        world, st[1], st[3], st[2], st[4], st[5], st.sum(), 2.93 / world, 4.03 / world), flush=True)
    h.close()
