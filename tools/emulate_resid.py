#!/usr/bin/env python
"""numpy emulation of the CURRENT CUDA arithmetic (residual form, integer offsets, FP32 pair maths with FMA
chains, two-level FP32 group sums, FP64 beyond) with an exactly-rounded exp2 in place of MUFU.EX2.
CPU microscope: which part of the sigma2 error is arithmetic and which is MUFU."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import cpd_oracle as orc
import residual_form as rf

LOG2E = 1.4426950408889634
f32 = np.float32


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def chain(ai, bj, no):      # t' = u - o  with the kernel's FMA chain (ai: (k,3) f32, bj: (3,) f32, no: (k,) f32 = -o)
    d = (ai - bj).astype(f32)
    t = fma32(d[:, 0], d[:, 0], no)
    t = fma32(d[:, 1], d[:, 1], t)
    return fma32(d[:, 2], d[:, 2], t), d


def ex2(t):
    return np.exp2(-t.astype(np.float64)).astype(f32)


def estep(ts, tgt, sigma2, w, cx, sub=64, grp=8, acc64=False):
    m, n = ts.shape[0], tgt.shape[0]
    sk = np.sqrt(LOG2E / (2.0 * sigma2))
    a = (sk * (ts - cx)).astype(f32)
    b = (sk * (tgt - cx)).astype(f32)
    at = np.float64 if acc64 else f32
    # pass 1: offsets seeded from the first 8 sources
    d8 = ((b[:, None, :] - a[None, :8, :]).astype(np.float64) ** 2).sum(-1).min(1)
    o = np.minimum(2.0 ** 20, np.floor(d8)).astype(f32)
    S = np.zeros(n); SU = np.zeros(n)
    for j0 in range(0, m, sub):
        js = list(range(j0, min(j0 + sub, m)))
        while True:
            Sc = np.zeros(n, dtype=at); Uc = np.zeros(n, dtype=at)
            for g0 in range(0, len(js), grp or len(js)):
                gs = np.zeros(n, dtype=at); gu = np.zeros(n, dtype=at)
                for j in js[g0:g0 + (grp or len(js))]:
                    t, _ = chain(b, a[j], -o)
                    e = ex2(t)
                    if acc64:
                        gs += e.astype(np.float64); gu += e.astype(np.float64) * t.astype(np.float64)
                    else:
                        gs = (gs + e).astype(f32); gu = fma32(e, t, gu)
                Sc = (Sc + gs).astype(at); Uc = (Uc + gu).astype(at)
            bad = ~(Sc < 2.0 ** 100)
            if not bad.any():
                break
            cm = np.min(((b[:, None, :] - a[None, js, :]).astype(np.float64) ** 2).sum(-1), axis=1)
            on = np.minimum(o, np.floor(cm).astype(f32))
            sh = np.maximum(on - o, -4000).astype(int)
            S = np.ldexp(S, sh); SU = np.ldexp(SU, sh); o = on
        S += Sc.astype(np.float64)
        SU += Uc.astype(np.float64) + o.astype(np.float64) * Sc.astype(np.float64)
    with np.errstate(divide="ignore"):
        log2S = np.log2(S) - o.astype(np.float64)
    dim = ts.shape[1]
    c = (2 * np.pi * sigma2) ** (dim / 2) * (w / (1 - w) * m / n) if w > 0 else 0.0
    dead = ~(log2S >= -1075.0)
    if c > 0:
        lc = np.log2(c); hi, lo = np.maximum(log2S, lc), np.minimum(log2S, lc)
        L = hi + np.log2(1 + np.exp2(lo - hi)); pt1 = np.exp2(log2S - L)
    else:
        L = log2S.copy(); pt1 = np.ones(n)
    pt1[dead] = 0
    rn = np.exp2(-(L + o.astype(np.float64))).astype(f32)
    rn[dead] = 0
    no = np.where(dead, np.inf, -o).astype(f32)
    srr = float((SU * rn.astype(np.float64))[~dead].sum())
    # pass 2
    A = np.zeros((m, 4))
    for j0 in range(0, n, sub):
        js = list(range(j0, min(j0 + sub, n)))
        s = np.zeros((m, 4), dtype=at)
        for g0 in range(0, len(js), grp or len(js)):
            g = np.zeros((m, 4), dtype=at)
            for j in js[g0:g0 + (grp or len(js))]:
                t, d = chain(a, b[j], np.full(m, no[j], dtype=f32))
                pr = (ex2(t).astype(np.float64) * np.float64(rn[j])).astype(f32)
                if acc64:
                    g[:, 0] += pr; g[:, 1:] += pr[:, None].astype(np.float64) * d.astype(np.float64)
                else:
                    g[:, 0] = (g[:, 0] + pr).astype(f32)
                    for k in range(3):
                        g[:, 1 + k] = fma32(pr, d[:, k], g[:, 1 + k])
            s = (s + g).astype(at)
        A += s.astype(np.float64)
    p1 = A[:, 0]
    v = -A[:, 1:] / sk
    return pt1, p1, v, srr / sk ** 2


def registration(src, tgt, iters, w=0.0, **kw):
    s2 = orc.sigma2_init_exact(src, tgt)
    cx, cy = tgt.mean(0), src.mean(0)
    A, t, scale = np.identity(3), np.zeros(3), 1.0
    yt = src - cy
    for _ in range(iters):
        Aold = scale * A
        ts = src.dot(Aold.T) + t
        pt1, p1, v, srr = estep(ts, tgt, s2, w, cx, **kw)
        mom = {"Np": float(p1.sum()), "Sy": yt.T.dot(p1), "C": (yt.T * p1).dot(yt), "V1": v.sum(0), "VY": v.T.dot(yt),
               "Srr": srr, "cy": cy}
        A, t, scale, s2, q = rf.mstep_residual(mom, Aold, t, 3)
    return s2


if __name__ == "__main__":
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "synthetic1500.npz"))
    ref = float(g["rigid20_sigma2"])
    for name, kw in [("sub64 grp8", {}), ("sub64 flat", {"grp": 0}), ("acc64 (exact sums)", {"acc64": True})]:
        s2 = registration(g["source"], g["target"], 20, **kw)
        print("%-22s rel.err %+.3e" % (name, (s2 - ref) / ref), flush=True)


def morton_order(pts, bits=10):
    lo, hi = pts.min(0), pts.max(0)
    q = np.clip(((pts - lo) / (hi - lo).max() * (2 ** bits - 1)).astype(np.int64), 0, 2 ** bits - 1)
    code = np.zeros(len(pts), dtype=np.int64)
    for b in range(bits):
        for a in range(pts.shape[1]):
            code |= ((q[:, a] >> b) & 1) << (pts.shape[1] * b + a)
    return np.argsort(code, kind="stable")
