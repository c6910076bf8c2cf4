"""TEST INFRASTRUCTURE: run a few library paths under the CPU emulation and print a digest of the results.
tests/test_emulated_library.py runs this under CPD_EMU_SCHED=rr / reverse / random and requires identical digests: a kernel
whose output depends on the order in which a block's threads run is missing a barrier."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpd_oracle as orc  # noqa: E402
from probreg_b200 import _cabi  # noqa: E402

_cabi._lib = _cabi._load(sys.argv[1])
from probreg_b200 import bcpd, cpd  # noqa: E402

h = hashlib.sha256()
src, tgt = orc.synthetic_pair(900)
r = cpd.registration_cpd(src, tgt, maxiter=4, tol=-1.0, w=0.1)
h.update(r.transformation.rot.tobytes() + np.float64(r.sigma2).tobytes())
ts = orc.apply_rigid(src, orc.rot_z(30.0), np.array([0.1, -0.2, 0.3]))
hd = _cabi.Handle(3)
hd.set_source(ts)
hd.set_target(tgt)
for arr in hd.estep(ts, 5e-5, 0.0)[:3]:                       # culling instantiations
    h.update(arr.tobytes())
lr = cpd.NonRigidCPD(src[:200], low_rank=24)
res = lr.registration(tgt[:230], maxiter=2, tol=-1.0)
h.update(res.transformation.w.tobytes() + lr.moved_source().tobytes())
es = bcpd.CombinedBCPD(src[:300]).expectation_step(src[:300], tgt[:280], 1.0, 1.0 / 300, np.ones(300), 0.02, 0.1)
h.update(es.nu.tobytes() + es.px.tobytes())
print(h.hexdigest())
