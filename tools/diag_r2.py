#!/usr/bin/env python
"""Round-2 hardware diagnostics of the three tests that failed on the B200 in round 1 (convergence traces, not pass/fail)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import cpd_oracle as orc
from probreg_b200 import _cabi
from probreg_b200.synthetic import synthetic_pair

F = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])


def affine_trace(n, iters, every):
    src, tgt = synthetic_pair(n, "affine")
    lin_true = orc.rot_z(30.0).dot(np.diag([1.1, 0.9, 1.05])); lin_true[0, 1] += 0.05
    h = _cabi.Handle(3); h.set_source(src); h.set_target(tgt)
    s2 = h.sigma2_init()
    h.set_state(_cabi.TF_AFFINE, True, 0.0, np.identity(3), np.zeros(3), 1.0, s2, 1.0 + n * 1.5 * np.log(s2))
    print("affine n=%d sigma2_0=%.5g" % (n, s2))
    qp = None
    for it in range(1, iters + 1):
        lin, t, sc, s2, q, npp = h.em_step()
        if it % every == 0 or it == 1:
            print("  it %4d sigma2 %.6e |B-B*|max %.4e |t-t*|max %.4e dq %.3e" % (it, s2, np.abs(lin - lin_true).max(),
                  np.abs(t - np.array([0.1, -0.2, 0.3])).max(), abs(q - qp) if qp is not None else float("nan")))
        qp = q


def deformed(m, seed=9):
    src, _ = synthetic_pair(m)
    tgt = src + 0.03 * np.sin(2 * np.pi * src.dot(F)) + 0.002 * np.random.default_rng(seed).standard_normal(src.shape)
    return src, tgt


def nonrigid_trace(m, iters, every, rank):
    src, tgt = deformed(m)
    base = np.sqrt(((src - tgt) ** 2).sum(1)).mean()
    h = _cabi.Handle(3); h.set_source(src); h.set_target(tgt)
    s2 = h.sigma2_init()
    t0 = time.perf_counter()
    if rank:
        h.nonrigid_lowrank_begin(2.0, 2.0, s2, 0.0, rank, 2, 0)
    else:
        h.nonrigid_begin(2.0, 2.0, s2, 0.0)
    print("nonrigid m=%d rank=%s sigma2_0=%.5g base residual %.5f" % (m, rank or "dense", s2, base))
    out = []
    for it in range(1, iters + 1):
        s2 = h.nonrigid_step()
        if it % every == 0 or it == 1:
            moved = h.nonrigid_moved()
            res = np.sqrt(((moved - tgt) ** 2).sum(1)).mean()
            out.append((it, s2, res))
            print("  it %4d sigma2 %.6e residual %.5f (%.3f of base)  [%.1f s]" % (it, s2, res, res / base, time.perf_counter() - t0))
    return out


which = sys.argv[1:] or ["affine", "nonrigid"]
if "affine" in which:
    affine_trace(3000, 300, 25)
    affine_trace(30000, 300, 25)
    affine_trace(250000, 400, 25)
if "nonrigid" in which:
    a = nonrigid_trace(12000, 100, 10, 0)
    b = nonrigid_trace(12000, 100, 10, 200)
    print("dense vs low-rank at 12000: max |sigma2 diff| rel %.3e, max residual diff %.3e" % (
        max(abs(x[1] - y[1]) / x[1] for x, y in zip(a, b)), max(abs(x[2] - y[2]) for x, y in zip(a, b))))
    nonrigid_trace(50000, 150, 10, 200)
