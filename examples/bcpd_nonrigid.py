#!/usr/bin/env python
"""BCPD (similarity + non-rigid) with the E-step on the GPU -- counterpart of the reference's examples/bcpd_nonrigid.py on a
synthetic pair (no open3d / transforms3d needed).  The M-step is dense M x M algebra on the host, as in the reference, so
keep the point count in the low thousands.   usage: python examples/bcpd_nonrigid.py [points]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probreg_b200 import bcpd
from probreg_b200.synthetic import synthetic_pair

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
source, target = synthetic_pair(n)
f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
target = target + 0.01 * np.sin(2 * np.pi * target.dot(f))
# (the reference's BCPD is fragile on unnormalised clouds -- its sigma2 update can overshoot below zero after a few more
#  iterations on this pair, with either implementation; tests/golden/bcpd.npz pins the first five against the reference)
tf_param = bcpd.registration_bcpd(source, target, w=0.05, maxiter=iters, tol=-1.0)
ang = np.rad2deg(np.arctan2(tf_param.rigid_trans.rot[1, 0], tf_param.rigid_trans.rot[0, 0]))
print("result: rotation about z %.2f deg, scale %.4f, t %s" % (ang, tf_param.rigid_trans.scale, tf_param.rigid_trans.t))
print("mean |v| of the non-rigid part: %.4f" % np.linalg.norm(tf_param.v, axis=1).mean())
