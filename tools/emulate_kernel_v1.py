#!/usr/bin/env python
"""[historical: the FIRST formulation (px sums, real-valued offsets); see emulate_resid.py for the current one]
numpy emulation of the CUDA E-step arithmetic (FP32 pair math, FMA chains, 64-point sub-chunks,
lazy offset) for small clouds -- a CPU microscope for precision questions, not product code.
Switches: acc64 (accumulate sub-chunks in float64), coord64 (no FP32 rounding of coordinates),
pair64 (pair maths in float64), sub (sub-chunk length)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cpd_oracle as orc

LOG2E = 1.4426950408889634
f32 = np.float32


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def ex2_32(t):
    return np.exp2(t.astype(np.float64)).astype(f32)


def estep_emul(ts, tgt, sigma2, w, cx, sub=64, acc64=False, coord64=False, pair64=False, grp=0):
    m, n = ts.shape[0], tgt.shape[0]
    sk = np.sqrt(LOG2E / (2.0 * sigma2))
    ct = np.float64 if coord64 else f32
    a = (sk * (ts - cx)).astype(ct)
    b = (sk * (tgt - cx)).astype(ct)
    pt = np.float64 if pair64 else f32
    at = np.float64 if acc64 else f32

    def tval(ai, bj, add):  # add - |ai-bj|^2 with the kernel's FMA chain; ai: (k,3) bj: (3,)
        if pair64:
            d = ai.astype(np.float64) - bj.astype(np.float64)
            return add.astype(np.float64) - (d * d).sum(1)
        d = (ai - bj).astype(f32)
        t = fma32(-d[:, 0], d[:, 0], add.astype(f32))
        t = fma32(-d[:, 1], d[:, 1], t)
        return fma32(-d[:, 2], d[:, 2], t)

    # pass 1: i = targets, j = sources
    o = np.full(n, 2.0 ** 20, dtype=f32)
    S = np.zeros(n)
    for j0 in range(0, m, sub):
        js = range(j0, min(j0 + sub, m))
        Sc = np.zeros(n, dtype=at)
        G = np.zeros(n, dtype=at)
        for cnt, j in enumerate(js):
            e = np.exp2(tval(b, a[j], o).astype(np.float64)).astype(pt if pair64 else f32)
            if grp:
                G = (G + e.astype(at)).astype(at)
                if (cnt + 1) % grp == 0 or cnt == len(js) - 1:
                    Sc = (Sc + G).astype(at); G = np.zeros(n, dtype=at)
            else:
                Sc = (Sc + e.astype(at)).astype(at)
        bad = ~(Sc < 2.0 ** 100)
        if bad.any():
            cm = np.full(n, 3e38)
            for j in js:
                d = (b - a[j]).astype(np.float64)
                cm = np.minimum(cm, (d * d).sum(1))
            on = np.minimum(o, np.floor(cm).astype(f32))
            S = np.ldexp(S, np.maximum(on - o, -4000).astype(int))
            o = on
            Sc = np.zeros(n, dtype=at)
            for j in js:
                e = np.exp2(tval(b, a[j], o).astype(np.float64)).astype(pt if pair64 else f32)
                Sc = (Sc + e.astype(at)).astype(at)
        S = S + Sc.astype(np.float64)
    with np.errstate(divide="ignore"):
        log2S = np.log2(S) - o.astype(np.float64)
    dim = ts.shape[1]
    c = 0.0
    if w > 0:
        c = (2 * np.pi * sigma2) ** (dim / 2) * (w / (1 - w) * m / n)
    dead = ~(log2S >= -1075.0)
    if c > 0:
        lc = np.log2(c)
        hi, lo = np.maximum(log2S, lc), np.minimum(log2S, lc)
        L = hi + np.log2(1 + np.exp2(lo - hi))
        pt1 = np.exp2(log2S - L)
    else:
        L = log2S.copy()
        pt1 = np.ones(n)
    pt1[dead] = 0.0
    Lhi = np.rint(L * 1024) / 1024
    g = np.exp2(Lhi - L)
    negLhi = np.where(dead, -np.inf, -Lhi).astype(ct)
    q1 = np.c_[g, g[:, None] * b.astype(np.float64)].astype(ct)
    q1[dead] = 0
    # pass 2: i = sources, j = targets
    A = np.zeros((m, 4))
    for j0 in range(0, n, sub):
        s = np.zeros((m, 4), dtype=at)
        G = np.zeros((m, 4), dtype=at)
        jl = list(range(j0, min(j0 + sub, n)))
        for cnt, j in enumerate(jl):
            p = np.exp2(tval(a, b[j], np.full(m, negLhi[j], dtype=ct)).astype(np.float64))
            if not pair64:
                p = p.astype(f32)
            if acc64 or pair64:
                s = (s.astype(np.float64) + p[:, None].astype(np.float64) * q1[j][None, :].astype(np.float64)).astype(at)
            elif grp:
                for k in range(4):
                    G[:, k] = fma32(p, np.full(m, q1[j, k], dtype=f32), G[:, k])
                if (cnt + 1) % grp == 0 or cnt == len(jl) - 1:
                    s = (s + G).astype(f32); G = np.zeros((m, 4), dtype=at)
            else:
                for k in range(4):
                    s[:, k] = fma32(p, np.full(m, q1[j, k], dtype=f32), s[:, k])
        A += s.astype(np.float64)
    p1 = A[:, 0]
    px = A[:, 1:] / sk + cx * p1[:, None]
    return orc.Estep(pt1, p1, px, float(p1.sum()))


def registration_emul(src, tgt, iters, w=0.0, **kw):
    s2 = orc.sigma2_init_exact(src, tgt)
    cx = tgt.mean(0)
    params = (np.identity(3), np.zeros(3), 1.0)
    for _ in range(iters):
        ts = orc.apply_rigid(src, *params)
        es = estep_emul(ts, tgt, s2, w, cx, **kw)
        r = orc.mstep_rigid(src, tgt, es)
        params, s2 = r.params, r.sigma2
    return r


if __name__ == "__main__":
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bunny.npz"))
    src, tgt = g["source"], g["target"]
    ref = float(g["rigid10_sigma2"])
    for name, kw in [("kernel arithmetic", {}), ("acc64", {"acc64": True}), ("coord64", {"coord64": True}),
                     ("pair64", {"pair64": True, "coord64": True, "acc64": True}), ("sub=16", {"sub": 16})]:
        r = registration_emul(src, tgt, 10, **kw)
        print("%-20s sigma2=%.12e  rel.err=%+.3e" % (name, r.sigma2, (r.sigma2 - ref) / ref))
