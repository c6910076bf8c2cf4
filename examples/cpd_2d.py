#!/usr/bin/env python
"""Affine and non-rigid CPD on the 2-D fish -- counterpart of the reference's examples/cpd_affine2d.py and cpd_nonrigid2d.py (which
plot every iteration with matplotlib; here the callback counts).  Needs the reference's fish_source.txt / fish_target.txt
(see utils.reference_file); without them a synthetic 2-D pair is used."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import utils  # noqa: E402
from probreg_b200 import cpd  # noqa: E402

fs, ft = utils.reference_file("fish_source.txt"), utils.reference_file("fish_target.txt")
if fs and ft:
    source, target = utils.prepare_source_and_target_nonrigid_2d(fs, ft)
else:
    rng = np.random.default_rng(0)
    source = rng.random((120, 2))
    target = source.dot(np.array([[1.1, 0.2], [-0.1, 0.9]]).T) + 0.05 * np.sin(4.0 * source[:, ::-1]) + 0.1
for name, kw in (("affine", {}), ("nonrigid", {"beta": 2.0, "lmd": 2.0})):
    seen = []
    res = cpd.registration_cpd(source, target, name, callbacks=[lambda t: seen.append(1)], **kw)
    moved = res.transformation.transform(source)
    gap = np.mean([np.linalg.norm(target - p, axis=1).min() for p in moved])
    print("%-9s %2d iterations, sigma2 %.3e, mean distance to the nearest target point %.4f" % (name, len(seen), res.sigma2, gap))
