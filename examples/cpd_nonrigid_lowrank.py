#!/usr/bin/env python
"""Non-rigid CPD at a size the dense M x M solve cannot reach (BASELINE configuration 5: N = M = 50k, rank 200).
Counterpart of the reference's examples/cpd_nonrigid3d_cuda.py (which builds the dense G and is limited to a few thousand
points).   usage: python examples/cpd_nonrigid_lowrank.py [points] [rank]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probreg_b200 import cpd
from probreg_b200.synthetic import synthetic_pair

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 200
source, _ = synthetic_pair(n)
f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
target = source + 0.03 * np.sin(2 * np.pi * source.dot(f))          # a smooth displacement field
reg = cpd.NonRigidCPD(source, beta=2.0, lmd=2.0, low_rank=rank)
t0 = time.perf_counter()
res = reg.registration(target, maxiter=60, tol=1e-7)      # q is sigma2 itself for the non-rigid family (cpd.py:303)
dt = time.perf_counter() - t0
moved = reg.moved_source()
print("M = N = %d, rank %d: %.2f s, sigma2 = %.3e" % (n, rank, dt, res.sigma2))
print("mean distance to the target: before %.4f, after %.4f" % (
    np.linalg.norm(source - target, axis=1).mean(), np.linalg.norm(moved - target, axis=1).mean()))
