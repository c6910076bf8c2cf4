#!/usr/bin/env python
"""Kernel-level effect of the exact culling: pass-1 / pass-2 times at decreasing sigma2, culling on vs off."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair, rot_z

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
src, tgt = synthetic_pair(n)
for no_cull in (True, False):
    os.environ["CPD_B200_NO_CULL"] = "1" if no_cull else "0"
    h = _cabi.Handle(3)
    h.set_source(src); h.set_target(tgt)
    for s2 in (1e-2, 1e-3, 3e-4, 1e-4, 1e-5):
        h.set_state(_cabi.TF_RIGID, True, 0.0, rot_z(30.0), np.array([0.1, -0.2, 0.3]), 1.0, s2, 0.0)
        h.em_step()
        h.set_state(_cabi.TF_RIGID, True, 0.0, rot_z(30.0), np.array([0.1, -0.2, 0.3]), 1.0, s2, 0.0)
        h.set_profiling(True)
        out = h.em_step()
        st = h.stage_times()
        h.set_profiling(False)
        print("cull %-3s sigma2 %.0e: pass1 %.3f ms pass2 %.3f ms  -> sigma2' %.6e" % ("off" if no_cull else "on", s2, st[1], st[3], out[3]), flush=True)
    h.close()
