"""Bayesian Coherent Point Drift -- the API surface of ``probreg.bcpd`` with the E-step on the B200.

SURVEY section 8(f) row 3: ``BayesianCoherentPointDrift.expectation_step`` (probreg/bcpd.py:53-72) is the CPD E-step with a
weight per source, and runs in the same two sm_100a passes (``cpd_bcpd_estep``: a template parameter of the pair kernels;
the M x N matrix the reference builds is never formed).  The M-step of ``CombinedBCPD`` (bcpd.py:127-156) is dense M x M
linear algebra on the host in the reference (two matrix inverses per iteration) and is outside the accelerated path: it
is restated here in numpy so that ``registration_bcpd`` runs end to end with the reference's signature and results.
"""
import abc
from collections import namedtuple

import numpy as np
import scipy.special as spsp
from scipy.spatial import cKDTree

from . import _cabi
from . import math_utils as mu
from . import transformation as tf
from .cpd import _points
from .log import log

EstepResult = namedtuple("EstepResult", ["nu_d", "nu", "n_p", "px", "x_hat"])
MstepResult = namedtuple("MstepResult", ["transformation", "u_hat", "sigma_mat", "alpha", "sigma2"])


class BayesianCoherentPointDrift(abc.ABC):
    """EM driver of probreg/bcpd.py:31-101.  ``source``: (M, D) array or None; ``device``: CUDA ordinal (extension)."""

    def __init__(self, source=None, device=0):
        self._source = None if source is None else _points(source)
        self._tf_type = None
        self._callbacks = []
        self._device = device
        self._h = None

    def set_source(self, source):
        self._source = _points(source)

    def set_callbacks(self, callbacks):
        self._callbacks.extend(callbacks)

    @abc.abstractmethod
    def _initialize(self, target):
        return MstepResult(None, None, None, None, None)

    def expectation_step(self, t_source, target, scale, alpha, sigma_mat, sigma2, w=0.0):
        """Expectation step for BCPD (probreg/bcpd.py:53-72) on the device.

        ``sigma_mat``: the M x M posterior covariance or just its diagonal (the only part bcpd.py:61 reads).
        Host arrays in, host arrays out: EstepResult(nu_d (N), nu (M), n_p, px (M, D), x_hat (M, D)).
        """
        t_source, target = np.asarray(t_source), np.asarray(target)
        assert t_source.ndim == 2 and target.ndim == 2, "source and target must have 2 dimensions."
        dim = t_source.shape[1]
        if self._h is None or self._h.dim != dim:
            self._h = _cabi.Handle(dim, device=self._device)
        sdiag = np.asarray(sigma_mat, dtype=np.float64)
        if sdiag.ndim == 2:
            sdiag = np.ascontiguousarray(np.diag(sdiag))
        self._h.set_source(t_source)
        self._h.set_target(target)
        alpha = np.broadcast_to(np.asarray(alpha, dtype=np.float64), (t_source.shape[0],))
        nu_d, nu, px, n_p = self._h.bcpd_estep(t_source, scale, alpha, sdiag, sigma2, w)
        with np.errstate(divide="ignore", invalid="ignore"):
            x_hat = px / nu[:, None]                                  # bcpd.py:70-71
        return EstepResult(nu_d, nu, n_p, px, x_hat)

    def maximization_step(self, target, estep_res, sigma2_p=None):
        return self._maximization_step(self._source, target, estep_res, sigma2_p)

    @staticmethod
    @abc.abstractmethod
    def _maximization_step(source, target, estep_res, sigma2_p=None):
        return None

    def registration(self, target, w=0.0, maxiter=50, tol=0.001):
        """probreg/bcpd.py:82-101: EM until the mean nearest-neighbour distance of the moved source stops changing."""
        assert not self._tf_type is None, "transformation type is None."
        target = _points(target)
        res = self._initialize(target)
        target_tree = cKDTree(target, leafsize=10)
        rmse = None
        for i in range(maxiter):
            t_source = res.transformation.transform(self._source)
            estep_res = self.expectation_step(t_source, target, res.transformation.rigid_trans.scale, res.alpha, res.sigma_mat,
                                              res.sigma2, w)
            res = self.maximization_step(target, res.transformation.rigid_trans, estep_res, res.sigma2)
            for c in self._callbacks:
                c(res.transformation)
            tmp_rmse = mu.compute_rmse(t_source, target_tree)
            log.debug("Iteration: {}, Criteria: {}".format(i, tmp_rmse))
            if not rmse is None and abs(rmse - tmp_rmse) < tol:
                break
            rmse = tmp_rmse
        return res.transformation


class CombinedBCPD(BayesianCoherentPointDrift):
    """Similarity + non-rigid BCPD (probreg/bcpd.py:104-156).  ``lmd``: weight of the motion-coherence prior;
    ``k``: Dirichlet concentration of the mixing weights; ``gamma``: scale of the initial sigma2."""

    def __init__(self, source=None, lmd=2.0, k=1.0e20, gamma=1.0, device=0):
        super(CombinedBCPD, self).__init__(source, device)
        self._tf_type = tf.CombinedTransformation
        self.lmd = lmd
        self.k = k
        self.gamma = gamma

    def _initialize(self, target):
        m, dim = self._source.shape
        self.gmat = mu.inverse_multiquadric_kernel(self._source, self._source, device=self._device)
        self.gmat_inv = np.linalg.inv(self.gmat)
        sigma2 = self.gamma * mu.squared_kernel_sum(self._source, target, device=self._device)
        return MstepResult(self._tf_type(np.identity(dim), np.zeros(dim)), None, np.identity(m), 1.0 / m, sigma2)

    def maximization_step(self, target, rigid_trans, estep_res, sigma2_p=None):
        return self._maximization_step(self._source, target, rigid_trans, estep_res, self.gmat_inv, self.lmd, self.k, sigma2_p)

    @staticmethod
    def _maximization_step(source, target, rigid_trans, estep_res, gmat_inv, lmd, k, sigma2_p=None):
        """probreg/bcpd.py:127-156 on host arrays (dense M x M algebra, as in the reference)."""
        nu_d, nu, n_p, px, x_hat = estep_res
        m, dim = source.shape
        ratio = rigid_trans.scale ** 2 / sigma2_p ** 2                                    # bcpd.py:131  (sic: sigma2 squared)
        sigma_mat = np.linalg.inv(lmd * gmat_inv + ratio * np.diag(nu))                   # bcpd.py:132-133
        back = rigid_trans.inverse().transform(x_hat) - source                            # bcpd.py:134
        v_hat = ratio * sigma_mat.dot(nu[:, None] * back)                                 # bcpd.py:135-137 (kron with I_D == per-coordinate)
        u_hat = source + v_hat
        alpha = np.exp(spsp.psi(k + nu) - spsp.psi(k * m + n_p))                          # bcpd.py:139
        x_m = nu.dot(x_hat) / n_p
        sigma2_m = nu.dot(np.diag(sigma_mat)) / n_p
        u_m = nu.dot(u_hat) / n_p
        u_c = u_hat - u_m
        s_xu = ((x_hat - x_m).T * nu).dot(u_c) / n_p                                      # bcpd.py:144
        s_uu = (u_c.T * nu).dot(u_c) / n_p + sigma2_m * np.identity(dim)                  # bcpd.py:145
        phi, _, psih = np.linalg.svd(s_xu, full_matrices=True)
        fix = np.ones(dim)
        fix[-1] = np.linalg.det(phi.dot(psih))
        rot = (phi * fix).dot(psih)                                                       # bcpd.py:146-149
        scale = np.trace(rot.dot(s_xu)) / np.trace(s_uu)
        t = x_m - scale * rot.dot(u_m)
        y_hat = rigid_trans.transform(source + v_hat)                                     # bcpd.py:153 (the PREVIOUS similarity)
        s1 = nu_d.dot(np.einsum("ij,ij->i", target, target))
        s2 = np.einsum("ij,ij->", px, y_hat)
        s3 = nu.dot(np.einsum("ij,ij->i", y_hat, y_hat))
        sigma2 = (s1 - 2.0 * s2 + s3) / (n_p * dim) + scale ** 2 * sigma2_m               # bcpd.py:157
        return MstepResult(tf.CombinedTransformation(rot, t, scale, v_hat), u_hat, sigma_mat, alpha, sigma2)


def registration_bcpd(source, target, w=0.0, maxiter=50, tol=0.001, callbacks=(), **kwargs):
    """BCPD registration with the signature of probreg/bcpd.py:159-185; returns the estimated Transformation."""
    bcpd = CombinedBCPD(_points(source), **kwargs)
    bcpd.set_callbacks(list(callbacks))
    return bcpd.registration(_points(target), w, maxiter, tol)
