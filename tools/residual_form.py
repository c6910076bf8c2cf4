#!/usr/bin/env python
"""numpy statement of the residual-form M-step (csrc/kernels.cuh: mstep_solve_residual) and of the
quantities the CUDA E-step hands it.  Used by tests/test_residual_algebra.py (CPU) to show that the
update form is algebraically the reference's M-step, and by tools/emulate_kernel.py."""
import numpy as np


def residual_moments(source, target, t_source, P):
    """From a dense P (M x N): the 24 moments the kernels accumulate (natural units)."""
    cy = source.mean(0)
    yt = source - cy
    p1 = P.sum(1)
    r = target[None, :, :] - t_source[:, None, :]          # x_n - T(y_m)
    v = (P[:, :, None] * r).sum(1)                          # v_m
    srr = float((P * (r ** 2).sum(-1)).sum())
    return {"Np": float(p1.sum()), "Sy": yt.T.dot(p1), "C": (yt.T * p1).dot(yt), "V1": v.sum(0), "VY": v.T.dot(yt),
            "Srr": srr, "cy": cy}


def mstep_residual(mom, A_old, t_old, dim, kind="rigid", update_scale=True):
    """A_old (D x D) = s R or B and t_old of the transform that produced t_source.  Returns (lin, t, scale, sigma2, q)."""
    Np, Sy, C, V1, VY, Srr, cy = (mom[k] for k in ("Np", "Sy", "C", "V1", "VY", "Srr", "cy"))
    muy = Sy / Np
    Y = C - np.outer(Sy, Sy) / Np
    Vc = VY - np.outer(V1, Sy) / Np
    Am = A_old.dot(Y) + Vc
    tr_atr = 0.0
    scale = 1.0
    if kind == "rigid":
        u, _, vh = np.linalg.svd(Am)
        fix = np.ones(dim)
        fix[-1] = np.linalg.det(u.dot(vh))
        lin = (u * fix).dot(vh)
        tr_atr = np.trace(Am.T.dot(lin))
        scale = tr_atr / np.trace(Y) if update_scale else 1.0
    else:
        lin = np.linalg.solve(Y.T, Am.T).T
    A_new = scale * lin
    dA = A_old - A_new
    Q = Srr + 2.0 * np.sum(dA * Vc) - V1.dot(V1) / Np + np.trace(dA.dot(Y).dot(dA.T))
    if kind == "rigid" and not update_scale:
        sigma2 = (Q + tr_atr) / (Np * dim)
    else:
        sigma2 = Q / (Np * dim)
    sigma2 = max(sigma2, float(np.finfo(np.float32).eps))
    q = Q / (2.0 * sigma2) + dim * Np * 0.5 * np.log(sigma2)
    mux = A_old.dot(cy + muy) + t_old + V1 / Np          # = sum_m px_m / Np
    t_new = mux - A_new.dot(cy + muy)
    return lin, t_new, scale, sigma2, q
