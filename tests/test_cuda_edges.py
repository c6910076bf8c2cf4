"""Edge cases of the CUDA path through the public API / C ABI: collisions (duplicate and coincident points),
degenerate and tiny clouds, extreme sigma2, handle reuse with other sizes, argument errors.  gpu-marked."""
import numpy as np
import pytest

from oracle import cpd_oracle as orc
from probreg_b200 import _cabi, cpd, math_utils

pytestmark = pytest.mark.gpu


def _close(es, ref, rtol=2e-5):
    np.testing.assert_allclose(es.pt1, ref.pt1, rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(es.p1, ref.p1, rtol=rtol, atol=1e-9)
    np.testing.assert_allclose(es.px, ref.px, rtol=rtol, atol=rtol * max(1e-12, np.abs(ref.px).max()))


def test_duplicate_points_and_exact_coincidence():
    """Collisions: every source appears three times, targets coincide exactly with sources (u == 0)."""
    base = np.random.default_rng(0).random((300, 3))
    src = np.ascontiguousarray(np.repeat(base, 3, axis=0))
    tgt = np.ascontiguousarray(np.r_[base, base[:50]])
    for s2, w in [(1e-2, 0.0), (1e-6, 0.1)]:
        es = cpd.RigidCPD(src).expectation_step(src, tgt, s2, w)
        _close(es, orc.expectation_step(src, tgt, s2, w))
    res = cpd.registration_cpd(src, tgt, maxiter=8, tol=-1.0)
    oref, _ = orc.registration(src, tgt, "rigid", maxiter=8, tol=-1.0)
    np.testing.assert_allclose(res.transformation.rot, oref.params[0], atol=1e-5)
    assert res.sigma2 == pytest.approx(oref.sigma2, rel=1e-6)


def test_identical_clouds_hit_the_sigma2_floor_like_the_reference():
    src = np.random.default_rng(1).random((200, 3))
    res = cpd.registration_cpd(src, src.copy())
    oref, it = orc.registration(src, src.copy(), "rigid")
    assert res.sigma2 == pytest.approx(oref.sigma2, rel=1e-6)
    assert oref.sigma2 == pytest.approx(float(np.finfo(np.float32).eps))         # cpd.py:189
    np.testing.assert_allclose(res.transformation.rot, np.identity(3), atol=1e-6)


def test_all_points_equal_and_collinear_clouds():
    one = np.tile(np.array([[0.3, -0.2, 0.9]]), (40, 1))
    tgt = np.random.default_rng(2).random((25, 3))
    es = cpd.RigidCPD(one).expectation_step(one, tgt, 0.05, 0.1)            # zero-extent cloud: Morton range is 0
    _close(es, orc.expectation_step(one, tgt, 0.05, 0.1))
    line = np.c_[np.linspace(0, 1, 64), np.zeros(64), np.zeros(64)]
    es = cpd.RigidCPD(line).expectation_step(line, line[::-1] + 0.01, 0.01, 0.0)
    _close(es, orc.expectation_step(line, np.ascontiguousarray(line[::-1] + 0.01), 0.01, 0.0))


@pytest.mark.parametrize("s2", [1e3, 1e-9])
def test_extreme_sigma2(s2):
    src, tgt = orc.synthetic_pair(700)
    es = cpd.RigidCPD(src).expectation_step(src, tgt, s2, 0.0)
    ref = orc.expectation_step(src, tgt, s2, 0.0)
    # sigma2 = 1e-9: every column underflows in float64 -> the reference returns all zeros (cpd.py:81); so must we
    assert (ref.pt1 == 0).sum() == (es.pt1 == 0).sum()
    _close(es, ref, rtol=2e-5)


def test_handle_reuse_with_other_sizes_and_families():
    rng = np.random.default_rng(3)
    r = cpd.RigidCPD(rng.random((50, 3)))
    for m, n in [(50, 70), (3000, 10), (10, 3000), (1200, 1300), (2, 2)]:
        src, tgt = rng.random((m, 3)), rng.random((n, 3))
        r.set_source(src)
        es = r.expectation_step(src, tgt, 0.02, 0.05)
        _close(es, orc.expectation_step(src, tgt, 0.02, 0.05))
        res = r.registration(tgt, maxiter=2, tol=-1.0)
        oref, _ = orc.registration(src, tgt, "rigid", maxiter=2, tol=-1.0)
        assert res.sigma2 == pytest.approx(oref.sigma2, rel=1e-6)


def test_same_sizes_new_data_on_one_handle_and_buffers_free_on_return():
    """set_source / set_target enqueue upload -> statistics -> frame -> sort without a host round trip (ingest_cloud): (i) the caller's
    buffers may be overwritten as soon as the calls return; (ii) a second pair of clouds of the SAME sizes on the same handle (same
    work lists, same captured graph, new centroid and extent taken from the device) gives what a fresh handle gives."""
    rng = np.random.default_rng(11)
    h = _cabi.Handle(3)
    for shift in (0.0, 7.5):
        src, tgt = orc.synthetic_pair(3000)
        src, tgt = src * (1.0 + shift) + shift, tgt * (1.0 + shift) + shift
        bs, bt = src.copy(), tgt.copy()
        h.set_source(bs)
        bs[...] = rng.random(bs.shape)                   # the upload has been consumed: scribbling must not matter
        h.set_target(bt)
        bt[...] = rng.random(bt.shape)
        s2 = h.sigma2_init()
        assert s2 == pytest.approx(orc.sigma2_init_exact(src, tgt), rel=1e-11)
        h.set_state(_cabi.TF_RIGID, True, 0.0, np.identity(3), np.zeros(3), 1.0, s2, 0.0)
        for _ in range(3):
            lin, t, scale, sigma2, q, n_p = h.em_step()
        ref, _ = orc.registration(src, tgt, "rigid", maxiter=3, tol=-1.0, sigma2_0=s2)
        assert sigma2 == pytest.approx(ref.sigma2, rel=1e-6)
        np.testing.assert_allclose(lin, ref.params[0], atol=1e-8)
        np.testing.assert_allclose(t, ref.params[1], atol=1e-7 * (1.0 + shift))
    h.close()


def test_inputs_are_not_modified_and_any_layout_is_accepted():
    src, tgt = orc.synthetic_pair(500)
    src_f = np.asfortranarray(src)                      # non C-contiguous, float32, lists: coerced like cv() (cpd.py:444)
    s0, t0 = src.copy(), tgt.copy()
    a = cpd.registration_cpd(src_f, tgt.astype(np.float32).astype(np.float64).tolist(), maxiter=3, tol=-1.0)
    b = cpd.registration_cpd(src, tgt.astype(np.float32).astype(np.float64), maxiter=3, tol=-1.0)
    assert a.sigma2 == b.sigma2
    assert np.array_equal(src, s0) and np.array_equal(tgt, t0)


def test_argument_errors():
    with pytest.raises(ValueError):
        _cabi.Handle(4)
    h = _cabi.Handle(3)
    with pytest.raises(_cabi.CpdError, match="source and target"):
        h.sigma2_init()
    h.set_source(np.zeros((5, 3)))
    h.set_target(np.ones((6, 3)))
    with pytest.raises(_cabi.CpdError, match="w must be"):
        h.set_state(_cabi.TF_RIGID, True, 1.0, np.identity(3), np.zeros(3), 1.0, 0.1, 0.0)
    with pytest.raises(_cabi.CpdError, match="sigma2 must be positive"):
        h.estep(np.zeros((5, 3)), 0.0, 0.0)
    with pytest.raises(_cabi.CpdError, match="cpd_set_state"):
        h.em_step()
    with pytest.raises(ValueError):
        h.estep(np.zeros((4, 3)), 0.1, 0.0)             # wrong number of rows
    with pytest.raises(ValueError):
        h.set_target(np.zeros((4, 2)))                  # wrong dimension
    with pytest.raises(ValueError, match="same dimensions"):
        math_utils.squared_kernel_sum(np.zeros((3, 3)), np.zeros((3, 2)))


def test_results_are_deterministic():
    src, tgt = orc.synthetic_pair(4000)
    a = cpd.registration_cpd(src, tgt, maxiter=6, tol=-1.0)
    b = cpd.registration_cpd(src, tgt, maxiter=6, tol=-1.0)
    assert a.sigma2 == b.sigma2 and a.q == b.q
    assert np.array_equal(a.transformation.rot, b.transformation.rot)


def test_gauss_transform_vs_direct():
    """reference tests/test_gauss_transform.py compares IFGT with the direct sum at 1e-4; this is the direct sum itself."""
    from probreg_b200 import gauss_transform as gt
    rng = np.random.default_rng(5)
    src, tgt = rng.random((777, 3)), rng.random((333, 3))
    w1 = rng.standard_normal(777)
    for h in (1.0, 0.5, 0.05):
        direct = np.array([np.dot(w1, np.exp(-np.sum((t - src) ** 2, axis=1) / h ** 2)) for t in tgt])     # gauss_transform.py:10-16
        np.testing.assert_allclose(gt.GaussTransform(src, h).compute(tgt, w1), direct, rtol=2e-5, atol=2e-5 * np.abs(direct).max())
    wk = rng.standard_normal((6, 777))
    out = gt.GaussTransform(src, 0.3).compute(tgt, wk)
    ref = np.exp(-((tgt[:, None, :] - src[None, :, :]) ** 2).sum(-1) / 0.09).dot(wk.T).T
    assert out.shape == (6, 333)
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-5 * np.abs(ref).max())
    s2, t2 = rng.random((50, 2)), rng.random((20, 2))
    ref2 = np.exp(-((t2[:, None, :] - s2[None, :, :]) ** 2).sum(-1) / 0.25).sum(1)
    np.testing.assert_allclose(gt.GaussTransform(s2, 0.5).compute(t2), ref2, rtol=2e-5)


def test_culling_is_bit_exact_and_faster():
    """Skipping (warp, stage) blocks whose every 2^-(u-o) flushes to zero must not change a single bit."""
    import os
    import time
    src, tgt = orc.synthetic_pair(30000)
    ts = orc.apply_rigid(src, orc.rot_z(30.0), np.array([0.1, -0.2, 0.3]))

    def run(no_cull):
        os.environ["CPD_B200_NO_CULL"] = "1" if no_cull else "0"
        try:
            h = _cabi.Handle(3)
        finally:
            os.environ.pop("CPD_B200_NO_CULL", None)
        h.set_source(ts)
        h.set_target(tgt)
        out = {}
        for s2 in (1e-5, 3e-4, 0.05):
            h.estep(ts, s2, 0.1)                       # warm
            t0 = time.perf_counter()
            out[s2] = h.estep(ts, s2, 0.1)
            out[(s2, "t")] = time.perf_counter() - t0
        h.set_state(_cabi.TF_RIGID, True, 0.0, np.identity(3), np.zeros(3), 1.0, 2e-4, 0.0)
        h.set_source(src)
        h.set_target(tgt)
        h.set_state(_cabi.TF_RIGID, True, 0.0, orc.rot_z(29.0), np.array([0.1, -0.2, 0.3]), 1.0, 2e-4, 0.0)
        out["run"] = h.em_run(6, -1.0)
        return out

    a, b = run(False), run(True)
    for s2 in (1e-5, 3e-4, 0.05):
        for x, y in zip(a[s2][:3], b[s2][:3]):
            assert np.array_equal(x, y), "culling changed the E-step at sigma2=%g" % s2
        assert a[s2][3] == b[s2][3]
    assert a["run"][3] == b["run"][3] and a["run"][4] == b["run"][4]          # sigma2, q after 6 iterations
    assert np.array_equal(a["run"][0], b["run"][0])
    print("E-step 30k x 30k, sigma2=1e-5: culled %.2f ms vs dense %.2f ms" % (a[(1e-5, "t")] * 1e3, b[(1e-5, "t")] * 1e3))
    # (kernel-level timings: tools/cull_probe.py; this call is dominated by the host copies)
