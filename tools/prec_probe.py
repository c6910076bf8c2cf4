#!/usr/bin/env python
"""sigma2 / rotation error vs the reference fixtures for every build/variants/lib_*.so (precision tuning)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from probreg_b200 import cpd
out = []
for fname, tag, kind, it, w, sk, tk in [("bunny.npz","rigid10","rigid",10,0.0,"source","target"),
        ("synthetic1500.npz","rigid20","rigid",20,0.0,"source","target"),
        ("synthetic1500.npz","rigid20_outl_w","rigid",20,0.2,"source","target_outl"),
        ("synthetic1500.npz","rigid30_outl_w0","rigid",30,0.0,"source","target_outl"),
        ("synthetic1500.npz","affine20","affine",20,0.0,"source_a","target_a")]:
    g = np.load(%r + "/tests/golden/" + fname)
    r = cpd.registration_cpd(g[sk], g[tk], kind, w=w, maxiter=it, tol=-1.0)
    ref = float(g[tag + "_sigma2"])
    out.append("%%s %%+.2e" %% (tag, (r.sigma2 - ref) / ref))
print("  ".join(out))
''' % (ROOT, ROOT)
libs = [os.path.join(ROOT, "probreg_b200", "libcpd_b200.so")] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "lib_*.so")))
for lib in libs:
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, CPD_B200_LIB=lib), capture_output=True, text=True)
    print("%-28s %s %s" % (os.path.basename(lib), r.stdout.strip(), r.stderr.strip()[-200:]))
