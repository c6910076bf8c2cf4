#!/usr/bin/env python
"""Rigid CPD on the bunny -- counterpart of the reference's examples/cpd_rigid.py (no open3d viewer: the callback prints).
Uses the reference's bunny.pcd when it can be found (see utils.reference_file), a synthetic pair otherwise."""
import logging
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import utils  # noqa: E402
from probreg_b200 import cpd  # noqa: E402
from probreg_b200.synthetic import synthetic_pair  # noqa: E402

logging.getLogger("probreg").setLevel(logging.INFO)
path = utils.reference_file("bunny.pcd")
if path:
    source, target = utils.prepare_source_and_target_rigid_3d(path, rng=np.random.default_rng(0))
else:
    source, target = synthetic_pair(2000)
seen = []
tf_param, sigma2, q = cpd.registration_cpd(source, target, callbacks=[lambda t: seen.append(t.scale)])       # the reference example's call
angle = np.rad2deg(np.arctan2(tf_param.rot[1, 0], tf_param.rot[0, 0]))
print("iterations: %d, rotation about z: %.3f deg, scale: %.5f, t: %s, sigma2: %.3e" % (len(seen), angle, tf_param.scale, tf_param.t, sigma2))
