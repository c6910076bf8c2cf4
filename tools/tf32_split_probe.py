#!/usr/bin/env python
"""Round-2 groundwork (CPU experiment, not product code): which tensor-core input format could carry the G X products of the
low-rank path (csrc/lowrank.cuh) without hurting parity?  G X is the one GEMM of this repository whose A operand is generated
on the fly; the CUDA-core kernel is bound by 200 FP32 FMAs per pair.  Operands are rounded the way the tensor cores see them
(TF32: 10 explicit mantissa bits, BF16: 7; round to nearest even), products accumulate in float32 like the MMA does, and the
result is compared with the float64 product -- for one-term, the usual 3-term (hi*hi + hi*lo + lo*hi) and, for BF16, 6-term
splits -- both on G X itself and on what matters downstream: the moved points of one low-rank M-step."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cpd_oracle as orc


def round_mantissa(a, bits):
    """float32 array rounded (nearest-even) to `bits` explicit mantissa bits."""
    if bits >= 23:
        return np.ascontiguousarray(a, dtype=np.float32)
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    u = u + ((1 << (drop - 1)) - 1) + ((u >> drop) & 1)
    u = (u >> drop) << drop
    return u.astype(np.uint32).view(np.float32)


def split(a, bits, terms):
    out, rest = [], np.asarray(a, dtype=np.float32)
    for _ in range(terms):
        hi = round_mantissa(rest, bits)
        out.append(hi)
        rest = (rest - hi).astype(np.float32)
    return out


def mm32(a, b):
    return (a.astype(np.float32) @ b.astype(np.float32)).astype(np.float32)      # float32 accumulation


def product(g, x, bits, scheme):
    gs, xs = split(g, bits, 3), split(x, bits, 3)
    if scheme == 1:
        pairs = [(0, 0)]
    elif scheme == 3:
        pairs = [(0, 0), (0, 1), (1, 0)]
    else:
        pairs = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]
    acc = np.zeros((g.shape[0], x.shape[1]), dtype=np.float32)
    for i, j in sorted(pairs, key=lambda p: -(p[0] + p[1])):                     # small terms first
        acc = acc + mm32(gs[i], xs[j])
    return acc


m, k, beta, lmd = 3000, 64, 2.0, 2.0
src, _ = orc.synthetic_pair(m)
f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
tgt = src + 0.03 * np.sin(2 * np.pi * src.dot(f))
g32 = orc.rbf_kernel_f32(src, src, beta)
g = g32.astype(np.float64)
es = orc.expectation_step(src, tgt, 0.01, 0.0)
q_ref, b_ref = orc.lowrank_factors(src, beta, k)
t_ref = orc.mstep_nonrigid_lowrank(src, tgt, es, 0.01, q_ref, b_ref, lmd).params[1]
x = q_ref.astype(np.float32)
exact = g @ q_ref
print("M = %d, K = %d, beta = %.1f;  |G X| max = %.3g" % (m, k, beta, np.abs(exact).max()))
print("%-28s %12s %14s %16s" % ("operand format", "max |dGX|/|GX|", "|dBc|/|Bc|", "max |dT| (extent 1)"))
for name, bits, scheme in [("float32 (CUDA cores)", 23, 1), ("TF32 x1", 10, 1), ("TF32 x3", 10, 3), ("BF16 x1", 7, 1), ("BF16 x3", 7, 3),
                           ("BF16 x6", 7, 6)]:
    gx = product(g32, x, bits, scheme).astype(np.float64)
    bc = q_ref.T @ gx
    bc = 0.5 * (bc + bc.T)
    t = orc.mstep_nonrigid_lowrank(src, tgt, es, 0.01, q_ref, bc, lmd).params[1]
    print("%-28s %12.2e %14.2e %16.2e" % (name, np.abs(gx - exact).max() / np.abs(exact).max(),
                                          np.linalg.norm(bc - b_ref) / np.linalg.norm(b_ref), np.abs(t - t_ref).max()))
