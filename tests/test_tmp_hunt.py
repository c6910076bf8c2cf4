"""Randomised (hypothesis) checks of the library under the CPU emulation: shapes around the tile / stage / sub-chunk
boundaries (64, 512, 1024), sigma2 over ten decades, outlier weights, duplicated points, far outliers, 2-D and 3-D --
each against the numpy oracle.  Sizes are small because the emulation runs every CUDA thread as a fiber."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import cpd_oracle as orc
from probreg_b200 import _cabi, cpd

SIZES = st.sampled_from([1, 2, 31, 63, 64, 65, 127, 255, 256, 257, 511, 512, 513, 640, 1023, 1024, 1025, 1100])
COMMON = dict(max_examples=4000, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow],
              derandomize=False)


def _close(es, ref, rtol):
    np.testing.assert_array_equal(es.pt1 == 0, ref.pt1 == 0)
    np.testing.assert_allclose(es.pt1, ref.pt1, rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(es.p1, ref.p1, rtol=rtol, atol=1e-9)
    np.testing.assert_allclose(es.px, ref.px, rtol=rtol, atol=rtol * max(1e-12, np.abs(ref.px).max()))


@settings(**COMMON)
@given(m=SIZES, n=SIZES, dim=st.sampled_from([2, 3]), log_s2=st.floats(-7.0, 2.0), w=st.sampled_from([0.0, 0.0, 0.05, 0.5, 0.9]),
       seed=st.integers(0, 10 ** 6), shift=st.sampled_from([0.0, 0.0, 100.0, -3000.0]), dup=st.booleans())
def test_estep_matches_oracle_on_random_inputs(emulated, m, n, dim, log_s2, w, seed, shift, dup):
    rng = np.random.default_rng(seed)
    src = rng.random((m, dim)) * rng.uniform(0.2, 3.0) + shift
    tgt = rng.random((n, dim)) * rng.uniform(0.2, 3.0) + shift + rng.uniform(-0.2, 0.2)
    if dup and m > 3 and n > 3:
        src[1] = src[0]                       # duplicated source
        tgt[2] = src[0]                       # a target exactly on a source
        tgt[-1] = tgt[0] + 40.0               # a far outlier
    s2 = 10.0 ** log_s2
    es = cpd.RigidCPD(src).expectation_step(src, tgt, s2, w)
    ref = orc.expectation_step(src, tgt, s2, w)
    # columns whose largest exponent sits in float64's denormal band are noisy in the REFERENCE itself: skip those draws
    d2min = ((src[:, None, :] - tgt[None, :, :]) ** 2).sum(-1).min(0) / (2 * s2)
    if ((d2min > 690.0) & (d2min < 760.0)).any():
        return
    # element-wise error of FP32 coordinates in the sigma-scaled frame ~ extent / sigma * 1e-7 (see test_cuda_parity)
    extent = max(np.ptp(src, axis=0).max(), np.ptp(tgt, axis=0).max(), 1e-9)
    # ... and the exponent itself is carried in FP32: a target whose NEAREST source is u = d^2 / 2 sigma^2 away sees its column
    # perturbed by ~1e-7 u (found by a long random run: one target 29 sigma from every source, 6e-5 relative)
    rtol = max(2e-5, 3e-6 * extent / np.sqrt(s2), 3e-7 * float(d2min.max()))
    _close(es, ref, rtol=rtol)
    assert es.n_p == pytest.approx(ref.n_p, rel=rtol, abs=1e-9)     # a sum of few pairs when sigma << spacing: no averaging


@settings(**dict(COMMON, max_examples=25))
@given(m=st.sampled_from([60, 300, 700]), n=st.sampled_from([50, 333, 900]), kind=st.sampled_from(["rigid", "rigid_noscale", "affine"]),
       w=st.sampled_from([0.0, 0.1]), seed=st.integers(0, 10 ** 6), dim=st.sampled_from([2, 3]))
def test_registration_matches_oracle_on_random_inputs(emulated, m, n, kind, w, seed, dim):
    rng = np.random.default_rng(seed)
    src = rng.random((m, dim)) * np.array([1.0, 0.6, 0.3])[:dim]
    ang = rng.uniform(-0.5, 0.5)
    rot = np.identity(dim)
    rot[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
    lin = rot if kind != "affine" else rot.dot(np.diag(rng.uniform(0.8, 1.2, dim)))
    tgt = (src[rng.integers(0, m, n)] + 0.01 * rng.standard_normal((n, dim))).dot(lin.T) + rng.uniform(-0.3, 0.3, dim)
    tf_type = "affine" if kind == "affine" else "rigid"
    kw = {"update_scale": False} if kind == "rigid_noscale" else {}
    res = cpd.registration_cpd(src, tgt, tf_type, w=w, maxiter=6, tol=-1.0, **kw)
    ref, _ = orc.registration(src, tgt, tf_type, w=w, maxiter=6, tol=-1.0, update_scale=(kind != "rigid_noscale"))
    assert res.sigma2 == pytest.approx(ref.sigma2, rel=1e-6)
    lin_got = res.transformation.b if tf_type == "affine" else res.transformation.rot
    np.testing.assert_allclose(lin_got, ref.params[0], atol=1e-5)
    np.testing.assert_allclose(res.transformation.t, ref.params[1], atol=1e-5)
    if tf_type == "rigid":
        assert res.transformation.scale == pytest.approx(ref.params[2], rel=1e-5)


@settings(**dict(COMMON, max_examples=12))
@given(m=st.sampled_from([700, 1500, 2600]), n=st.sampled_from([600, 1300, 2100]), log_s2=st.floats(-6.5, -3.3), w=st.sampled_from([0.0, 0.1]),
       seed=st.integers(0, 10 ** 6), clusters=st.booleans())
def test_culled_estep_is_bit_identical_on_random_inputs(emulated, monkeypatch, m, n, log_s2, w, seed, clusters):
    """The exact-culling instantiations against the dense ones: every output bit-identical (sigma small enough to cull)."""
    rng = np.random.default_rng(seed)
    if clusters:      # a few well separated blobs: many far (warp, stage) blocks
        centres = rng.uniform(-2.0, 2.0, (5, 3))
        src = centres[rng.integers(0, 5, m)] + 0.05 * rng.standard_normal((m, 3))
        tgt = centres[rng.integers(0, 5, n)] + 0.05 * rng.standard_normal((n, 3))
    else:
        src, tgt = rng.random((m, 3)), rng.random((n, 3))
    s2 = 10.0 ** log_s2
    outs = []
    for no_cull in ("0", "1"):
        monkeypatch.setenv("CPD_B200_NO_CULL", no_cull)
        h = _cabi.Handle(3)
        h.set_source(src)
        h.set_target(tgt)
        outs.append(h.estep(src, s2, w))
    for a, b in zip(outs[0][:3], outs[1][:3]):
        assert np.array_equal(a, b)
    assert outs[0][3] == outs[1][3]


@settings(**dict(COMMON, max_examples=40))
@given(m=SIZES, n=SIZES, dim=st.sampled_from([2, 3]), log_s2=st.floats(-5.0, 1.0), w=st.sampled_from([0.0, 0.1, 0.6]),
       seed=st.integers(0, 10 ** 6), conc=st.sampled_from([0.2, 1.0, 50.0]), smax=st.sampled_from([0.0, 1e-3, 1.0]))
def test_bcpd_estep_matches_oracle_on_random_inputs(emulated, m, n, dim, log_s2, w, seed, conc, smax):
    from probreg_b200 import bcpd

    rng = np.random.default_rng(seed)
    src = rng.random((m, dim)) * 2.0
    tgt = rng.random((n, dim)) * 2.0 + rng.uniform(-0.1, 0.1)
    alpha = rng.dirichlet(np.full(m, conc))
    sdiag = rng.uniform(0.0, smax, m) if smax > 0 else np.zeros(m)
    s2, scale = 10.0 ** log_s2, rng.uniform(0.7, 1.3)
    # columns whose largest term sits in float64's denormal band are noisy in the REFERENCE itself (and which of its
    # intermediate products underflows depends on the order it multiplies in): skip those draws
    x = ((src[:, None, :] - tgt[None, :, :]) ** 2).sum(-1) / (2 * s2)
    la = -np.log(np.maximum(alpha, 1e-300) * (1.0 - w)) + scale ** 2 * dim * sdiag / (2 * s2)
    lognorm = dim * 0.5 * np.log(2 * np.pi * s2)
    lo = (x + np.minimum(0.0, la[:, None] + lognorm)).min(0)
    hi = (x + np.maximum(0.0, la[:, None] + lognorm)).min(0)
    if ((hi > 690.0) & (lo < 760.0)).any():
        return
    es = bcpd.CombinedBCPD(src).expectation_step(src, tgt, scale, alpha, sdiag, s2, w)
    ref = orc.bcpd_expectation_step(src, tgt, scale, alpha, sdiag, s2, w)
    # log2 of the per-source weight spans up to scale^2 D smax / (2 s2) binades; FP32 keeps it to ~6e-8 of that span
    span = scale ** 2 * dim * smax / (2 * s2) * 1.4427 + 60.0
    rtol = max(5e-5, 3e-6 * 2.0 / np.sqrt(s2), 3e-7 * span)
    live = ref.nu_d > 1e-290
    np.testing.assert_array_equal(es.nu_d[~live] < 1e-280, True)
    np.testing.assert_allclose(es.nu_d[live], ref.nu_d[live], rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(es.nu, ref.nu, rtol=rtol, atol=1e-9 + rtol * ref.nu.max())
    np.testing.assert_allclose(es.px, ref.px, rtol=rtol, atol=1e-9 + rtol * np.abs(ref.px).max())


@settings(**dict(COMMON, max_examples=8))
@given(m=st.sampled_from([40, 97, 150]), beta=st.sampled_from([0.3, 1.0, 2.0, 5.0]), lmd=st.sampled_from([0.5, 2.0, 8.0]),
       w=st.sampled_from([0.0, 0.1]), rank_frac=st.sampled_from([0.15, 0.4, 1.0]), seed=st.integers(0, 10 ** 6), dim=st.sampled_from([2, 3]))
def test_nonrigid_dense_and_lowrank_on_random_inputs(emulated, m, beta, lmd, w, rank_frac, seed, dim):
    """Dense device loop vs the reference arithmetic; low-rank loop vs the reference arithmetic on the SAME G = Q Bc Q^T."""
    rng = np.random.default_rng(seed)
    src = rng.random((m, dim))
    n = m + int(rng.integers(-m // 4, m // 4))
    tgt = src[rng.integers(0, m, n)] + 0.04 * np.sin(5.0 * src[rng.integers(0, m, n)][:, ::-1]) + 0.003 * rng.standard_normal((n, dim))
    dense = cpd.NonRigidCPD(src, beta=beta, lmd=lmd)
    rd = dense.registration(tgt, w=w, maxiter=3, tol=-1.0)
    od, _ = orc.registration(src, tgt, "nonrigid", maxiter=3, tol=-1.0, beta=beta, lmd=lmd, w=w)
    g = orc.rbf_kernel_f32(src, src, beta)
    assert rd.sigma2 == pytest.approx(od.sigma2, rel=2e-5)
    np.testing.assert_allclose(dense.moved_source(), src + g.dot(od.params[0]), atol=5e-5)
    rank = max(2, int(rank_frac * m)) if m <= 97 else max(2, int(min(rank_frac, 0.4) * m))      # keep the emulation quick
    low = cpd.NonRigidCPD(src, beta=beta, lmd=lmd, low_rank=rank)
    rl = low.registration(tgt, w=w, maxiter=3, tol=-1.0)
    g_lr = rl.transformation.q.dot(rl.transformation.bcore).dot(rl.transformation.q.T)
    ol, _ = orc.registration(src, tgt, "nonrigid", maxiter=3, tol=-1.0, beta=beta, lmd=lmd, w=w, g=g_lr)
    assert rl.sigma2 == pytest.approx(ol.sigma2, rel=2e-5)
    np.testing.assert_allclose(low.moved_source(), src + g_lr.dot(ol.params[0]), atol=5e-5)
    if rank == m:                                            # nothing truncated: the two device loops agree as well
        assert rl.sigma2 == pytest.approx(rd.sigma2, rel=2e-5)
