#!/usr/bin/env python
"""[historical: the CPU prototype that preceded csrc/lowrank.cuh; the product solves the K x K system without the eigen-decomposition]
CPU prototype (not product code): non-rigid CPD with a rank-K G = Q L Q^T.

Checks, against the dense oracle at small M, (i) how fast the spectrum of the RBF Gram matrix decays for the reference's
default beta, (ii) that the Woodbury form of the M-step (cpd.py:296)
    (diag(p1) G + lmd s2 I) W = F,  F = px - diag(p1) Y
    W = (F - diag(p1) Q (lmd s2 L^-1 + Q^T diag(p1) Q)^-1 Q^T F) / (lmd s2)
reproduces the dense solve, and (iii) that a randomised range finder that only needs products G X (which the pair kernel can
form on the fly, never storing G) finds Q.  Findings are printed; DESIGN.md section 7 refers to them."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cpd_oracle as orc

m, beta, lmd = 2000, 2.0, 2.0
src, _ = orc.synthetic_pair(m)
f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
tgt = src + 0.03 * np.sin(2 * np.pi * src.dot(f))
G = orc.rbf_kernel_f32(src, src, beta).astype(np.float64)
ev = np.linalg.eigvalsh(G)[::-1]
print("spectrum of G (beta=%.1f, unit-size cloud): lambda_k / lambda_1 at k = 1, 5, 10, 20, 50, 100, 200:" % beta)
print("   ", ["%.1e" % (ev[k - 1] / ev[0]) for k in (1, 5, 10, 20, 50, 100, 200)])

es = orc.expectation_step(src, tgt, 0.02, 0.0)
dense = orc.mstep_nonrigid(src, tgt, es, 0.02, G.astype(np.float32), lmd)
F = es.px - (src.T * es.p1).T
rng = np.random.default_rng(0)
for K in (10, 20, 50, 200):
    # randomised range finder with 2 power iterations: only products G @ X are needed
    X = rng.standard_normal((m, K + 10))
    for _ in range(3):
        X, _ = np.linalg.qr(G.dot(X))
    B = X.T.dot(G.dot(X))
    lam, V = np.linalg.eigh(B)
    idx = np.argsort(lam)[::-1][:K]
    Q, L = X.dot(V[:, idx]), lam[idx]
    s = lmd * 0.02
    small = np.diag(s / L) + (Q.T * es.p1).dot(Q)
    W = (F - (es.p1[:, None] * Q).dot(np.linalg.solve(small, Q.T.dot(F)))) / s
    T_lr = src + Q.dot(L[:, None] * Q.T.dot(W))
    T_dense = src + G.dot(dense.params[0])
    print("K = %3d: max |T_lowrank - T_dense| = %.2e (extent 1), ||G - Q L Q^T||_2 / ||G||_2 = %.1e" % (
        K, np.abs(T_lr - T_dense).max(), np.linalg.norm(G - (Q * L).dot(Q.T), 2) / ev[0]))
