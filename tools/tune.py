#!/usr/bin/env python
"""Time the E-step stages of every build/variants/lib_*.so at the bench configuration (one subprocess each)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json, numpy as np
sys.path.insert(0, %r)
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair
n = int(sys.argv[1])
src, tgt = synthetic_pair(n)
h = _cabi.Handle(3); h.set_source(src); h.set_target(tgt)
s2 = h.sigma2_init()
h.set_state(_cabi.TF_RIGID, True, 0.0, np.identity(3), np.zeros(3), 1.0, s2, 1.0)
for _ in range(3): h.em_step(read=False)
h.set_profiling(True)
st = []
for _ in range(8):
    h.em_step(read=False); st.append(h.stage_times())
st = np.median(np.array(st), axis=0)
out = h.em_step()
print(json.dumps({"pass1": float(st[1]), "pass2": float(st[3]), "total": float(st.sum()), "sigma2": out[3]}))
''' % ROOT

n = sys.argv[1] if len(sys.argv) > 1 else "100000"
libs = [os.path.join(ROOT, "probreg_b200", "libcpd_b200.so")] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "lib_*.so")))
for lib in libs:
    env = dict(os.environ, CPD_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", CHILD, n], env=env, capture_output=True, text=True)
    print("%-40s %s %s" % (os.path.basename(lib), r.stdout.strip(), r.stderr.strip()[-300:]))
