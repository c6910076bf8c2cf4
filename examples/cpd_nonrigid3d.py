#!/usr/bin/env python
"""Non-rigid CPD on the face scans -- counterpart of the reference's examples/cpd_nonrigid3d_cuda.py (which needs cupy and open3d).
Uses the reference's face-x.txt / face-y.txt when they can be found, a synthetic deformation otherwise.  Add a rank to use the
low-rank G:  python examples/cpd_nonrigid3d.py [voxel_size] [low_rank]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import utils  # noqa: E402
from probreg_b200 import cpd  # noqa: E402
from probreg_b200.synthetic import synthetic_pair  # noqa: E402

voxel = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
rank = int(sys.argv[2]) if len(sys.argv) > 2 else None
fx, fy = utils.reference_file("face-x.txt"), utils.reference_file("face-y.txt")
if fx and fy:
    source, target = utils.prepare_source_and_target_nonrigid_3d(fx, fy, voxel_size=voxel)
else:
    source, _ = synthetic_pair(3000)
    target = source + 0.03 * np.sin(2 * np.pi * source.dot(np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])))
acpd = cpd.NonRigidCPD(source, low_rank=rank)
start = time.time()
tf_param, sigma2, _ = acpd.registration(target)
print("time: %.3f s, sigma2: %.4g, |w| max: %.4g" % (time.time() - start, sigma2, np.abs(tf_param.w).max()))
result = acpd.moved_source()
print("mean nearest-target distance: before %.4g, after %.4g" % (
    np.mean([np.linalg.norm(target - p, axis=1).min() for p in source[::25]]),
    np.mean([np.linalg.norm(target - p, axis=1).min() for p in result[::25]])))
