#!/usr/bin/env python
"""A few EM iterations at the bench configuration, for ncu:  ncu ... python tools/prof_step.py [points] [iters]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probreg_b200 import _cabi  # noqa: E402
from probreg_b200.synthetic import synthetic_pair  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
src, tgt = synthetic_pair(n)
h = _cabi.Handle(3)
h.set_source(src)
h.set_target(tgt)
s2 = h.sigma2_init()
h.set_state(_cabi.TF_RIGID, True, 0.0, np.identity(3), np.zeros(3), 1.0, s2, 1.0 + n * 1.5 * np.log(s2))
for _ in range(iters):
    out = h.em_step()
print("sigma2 after %d iterations: %.9g" % (iters, out[3]))
