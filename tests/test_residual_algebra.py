"""CPU check of the algebra behind the CUDA M-step: written as an UPDATE of the previous transform from
residual moments (tools/residual_form.py == csrc/kernels.cuh:mstep_solve_residual) it reproduces the
reference's M-step (oracle, pinned against probreg/cpd.py:160-192 / 219-244) to rounding."""
import os
import sys

import numpy as np
import pytest
from scipy.spatial.distance import cdist

from conftest import ROOT
from oracle import cpd_oracle as orc

sys.path.insert(0, os.path.join(ROOT, "tools"))
import residual_form as rf  # noqa: E402


def _dense_p(ts, tgt, s2, w):
    k = np.exp(-cdist(ts, tgt, "sqeuclidean") / (2 * s2))
    den = k.sum(0)
    den[den == 0] = np.finfo(np.float32).eps
    den += orc.outlier_constant(s2, w, ts.shape[0], tgt.shape[0], ts.shape[1])
    return k / den


@pytest.mark.parametrize("kind,update_scale", [("rigid", True), ("rigid", False), ("affine", True)])
@pytest.mark.parametrize("s2,w", [(0.05, 0.0), (2e-3, 0.2)])
def test_update_form_equals_reference_mstep(kind, update_scale, s2, w):
    src, tgt = orc.synthetic_pair(300, "affine" if kind == "affine" else "rigid")
    rng = np.random.default_rng(1)
    a_old = orc.rot_z(20.0).dot(np.diag([1.05, 0.95, 1.0])) if kind == "affine" else 0.9 * orc.rot_z(20.0)
    t_old = rng.standard_normal(3) * 0.1
    ts = src.dot(a_old.T) + t_old
    P = _dense_p(ts, tgt, s2, w)
    es = orc.Estep(P.sum(0), P.sum(1), P.dot(tgt), float(P.sum()))
    ref = orc.mstep_affine(src, tgt, es) if kind == "affine" else orc.mstep_rigid(src, tgt, es, update_scale)
    lin, t, scale, sigma2, q = rf.mstep_residual(rf.residual_moments(src, tgt, ts, P), a_old, t_old, 3, kind, update_scale)
    np.testing.assert_allclose(lin, ref.params[0], atol=1e-10)
    np.testing.assert_allclose(t, ref.params[1], atol=1e-10)
    if kind == "rigid":
        assert scale == pytest.approx(ref.params[2], rel=1e-10)
    assert sigma2 == pytest.approx(ref.sigma2, rel=1e-8)
    assert q == pytest.approx(ref.q, rel=1e-8)


def test_update_form_2d():
    g = np.load(os.path.join(ROOT, "tests", "golden", "nonrigid.npz"))
    src, tgt = g["fish_source"], g["fish_target"]
    a_old, t_old = np.identity(2), np.zeros(2)
    P = _dense_p(src, tgt, 0.05, 0.1)
    es = orc.Estep(P.sum(0), P.sum(1), P.dot(tgt), float(P.sum()))
    ref = orc.mstep_rigid(src, tgt, es)
    lin, t, scale, sigma2, q = rf.mstep_residual(rf.residual_moments(src, tgt, src, P), a_old, t_old, 2)
    np.testing.assert_allclose(lin, ref.params[0], atol=1e-10)
    np.testing.assert_allclose(t, ref.params[1], atol=1e-10)
    assert sigma2 == pytest.approx(ref.sigma2, rel=1e-9)
