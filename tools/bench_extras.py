#!/usr/bin/env python
"""Secondary measurements that ride along with bench.py's single-GPU run (key "extras" of its JSON line; never part of `value`).

bench.py starts this script as a SUBPROCESS after its own measurements (a fault here cannot touch the bench line) and merges
the one JSON object it prints.  Contents: BASELINE.json configuration 3 (affine, N = M = 250k) and 5 (non-rigid, low-rank K = 200,
N = M = 50k) timed with CUDA events on the library's stream, and three small hardware parity probes for the paths whose first
GPU run this is (low-rank vs dense at M = 2000, the BCPD E-step against the reference's own fixture, the device priors) -- each
guarded, errors reported as strings.   usage: python tools/bench_extras.py [--quick]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
QUICK = "--quick" in sys.argv


def guarded(fn):
    try:
        return fn()
    except Exception as e:                      # noqa: BLE001 -- reported, not raised: this is a side measurement
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def deformed(n, seed=9):
    from probreg_b200.synthetic import synthetic_pair

    src, _ = synthetic_pair(n)
    f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
    tgt = src + 0.03 * np.sin(2 * np.pi * src.dot(f)) + 0.002 * np.random.default_rng(seed).standard_normal(src.shape)
    return src, np.ascontiguousarray(tgt)


def affine_250k():
    from probreg_b200 import _cabi
    from probreg_b200.synthetic import synthetic_pair

    n = 20000 if QUICK else 250000
    src, tgt = synthetic_pair(n, "affine")
    h = _cabi.Handle(3)
    h.set_source(src)
    h.set_target(tgt)
    s2 = h.sigma2_init()
    h.set_state(_cabi.TF_AFFINE, True, 0.0, np.identity(3), np.zeros(3), 1.0, s2, 1.0 + n * 1.5 * np.log(s2))
    for _ in range(3):
        h.em_step(read=False)
    steps = 8
    for i in range(steps):
        h.event_record(2 * i)
        h.em_step(read=False)
        h.event_record(2 * i + 1)
    h.sync()
    ms = float(np.mean([h.event_elapsed(2 * i, 2 * i + 1) for i in range(steps)]))
    h.set_profiling(True)
    h.em_step(read=False)
    stages = h.stage_times()
    h.set_profiling(False)
    out = h.em_step()
    return {"workload": "affine CPD, synthetic 3-D N=M=%d (BASELINE config 3)" % n, "ms_per_iteration": ms, "it_per_s": 1e3 / ms,
            "gpair_per_s": 2.0 * n * n / (ms * 1e-3) / 1e9, "sigma2_after_13": out[3],
            "stage_ms": dict(zip(["pack", "pass1", "finalize1", "pass2", "finalize2", "moments_mstep"], [float(x) for x in stages]))}


def lowrank_50k():
    from probreg_b200 import _cabi

    n, rank = (4000, 64) if QUICK else (50000, 200)
    src, tgt = deformed(n)
    h = _cabi.Handle(3)
    h.set_source(src)
    h.set_target(tgt)
    s2 = h.sigma2_init()
    setup = {}
    for piters in (0, 2):                       # set-up = (piters + 2) products G X + (piters + 1) orthonormalisations + Bc
        h.sync()
        t0 = time.perf_counter()
        h.nonrigid_lowrank_begin(2.0, 2.0, s2, 0.0, rank, piters, 0)
        h.sync()
        setup["power_iters=%d" % piters] = (time.perf_counter() - t0) * 1e3
    setup_ms = setup["power_iters=2"]
    h.set_profiling(True)                       # the same set-up once more, phases timed with events inside the library
    h.nonrigid_lowrank_begin(2.0, 2.0, s2, 0.0, rank, 2, 0)
    phases = h.lowrank_setup_times()
    h.set_profiling(False)
    trace = [h.nonrigid_step()]
    l0 = h.launch_count()
    steps = 8
    for i in range(steps):
        h.event_record(2 * i)
        h.nonrigid_step()                       # returns sigma2: one small D2H + sync per iteration, as registration() does
        h.event_record(2 * i + 1)
    h.sync()
    ms = float(np.mean([h.event_elapsed(2 * i, 2 * i + 1) for i in range(steps)]))
    launches = (h.launch_count() - l0) / steps
    h.set_profiling(True)
    h.nonrigid_step()
    st = h.stage_times()
    h.set_profiling(False)
    stage_ms = {"estep": float(sum(st[:5])), "mstep_incl_KxK_solve": float(st[5])}
    trace.append(h.nonrigid_step())
    moved = h.nonrigid_moved()
    return {"workload": "non-rigid CPD, rank-%d G, synthetic 3-D N=M=%d, beta=lmd=2 (BASELINE config 5)" % (rank, n),
            "setup_ms": setup_ms, "setup_ms_by_power_iters": setup, "setup_phases_ms": phases, "ms_per_iteration": ms, "it_per_s": 1e3 / ms, "launches_per_iteration": launches, "stage_ms": stage_ms,
            "sigma2_first_and_10th": trace,
            "mean_residual_before_after": [float(np.linalg.norm(src - tgt, axis=1).mean()), float(np.linalg.norm(moved - tgt, axis=1).mean())]}


def bcpd_estep_100k():
    """The weighted (WGT) instantiations of pass 1 / pass 2 at the bench size: stage times of one BCPD E-step."""
    from probreg_b200 import _cabi
    from probreg_b200.synthetic import synthetic_pair

    n = 3000 if QUICK else 100000
    src, tgt = synthetic_pair(n)
    rng = np.random.default_rng(1)
    alpha, sdiag = rng.dirichlet(np.ones(n)), rng.uniform(0.0, 1e-3, n)
    h = _cabi.Handle(3)
    h.set_source(src)
    h.set_target(tgt)
    h.bcpd_estep(src, 1.0, alpha, sdiag, 0.02, 0.1)
    h.set_profiling(True)
    nu_d, nu, px, n_p = h.bcpd_estep(src, 1.0, alpha, sdiag, 0.02, 0.1)
    stages = h.stage_times()
    h.set_profiling(False)
    return {"workload": "BCPD E-step, synthetic 3-D N=M=%d, sigma2 0.02, w 0.1" % n,
            "stage_ms": dict(zip(["pack", "pass1_wgt", "finalize1", "pass2_wgt", "finalize2"], [float(x) for x in stages[:5]])),
            "conservation": {"n_p": n_p, "sum_nu_d": float(nu_d.sum()), "sum_nu": float(nu.sum())}}


def variants():
    """build/variants/lib_*.so (tools/variants.txt, built by __graft_entry__.build) next to the product library: E-step stage
    times at the bench shape and at the shard one rank of an 8-GPU run holds, and the low-rank set-up phases -- tuning data for
    the kernels the variants change (TMA stage size of the two passes; LR_COLS of lr_gram_apply_kernel)."""
    import glob

    from probreg_b200 import _cabi
    from probreg_b200.synthetic import synthetic_pair

    n, rank, nl = (3000, 48, 2000) if QUICK else (100000, 200, 50000)
    src, tgt = synthetic_pair(n)
    lsrc, ltgt = deformed(nl)
    out = {}
    libs = [("product", _cabi.lib())] + [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "lib_*.so")))]
    names = ["pack", "pass1", "finalize1", "pass2", "finalize2", "moments_mstep"]
    for name, lib in libs:
        def run(lib=lib):
            saved = _cabi._lib
            _cabi._lib = lib if not isinstance(lib, str) else _cabi._load(lib)
            try:
                h, hl = _cabi.Handle(3), _cabi.Handle(3)
            finally:
                _cabi._lib = saved
            res = {}
            for tag, lo, hi in (("full", 0, n), ("shard_1_of_8", 3 * (n // 8), 4 * (n // 8))):
                h.set_source(src)
                h.set_target(tgt[lo:hi], n_global=n, frame_origin=tgt.mean(axis=0))
                h.set_state(_cabi.TF_RIGID, True, 0.0, np.identity(3), np.zeros(3), 1.0, 0.05, 0.0)
                for _ in range(3):
                    h.em_step(read=False)
                h.set_profiling(True)
                st = []
                for _ in range(5):
                    h.flush_l2()
                    h.em_step(read=False)
                    st.append(h.stage_times())
                h.set_profiling(False)
                res["stage_ms_" + tag] = dict(zip(names, [float(x) for x in np.median(np.array(st), axis=0)]))
                res["sigma2_" + tag] = h.em_step()[3]
            hl.set_source(lsrc)
            hl.set_target(ltgt)
            s2 = hl.sigma2_init()
            hl.nonrigid_lowrank_begin(2.0, 2.0, s2, 0.0, rank, 0, 0)         # warm
            hl.set_profiling(True)
            hl.nonrigid_lowrank_begin(2.0, 2.0, s2, 0.0, rank, 0, 0)         # 2 products, 1 orthonormalisation
            ph = hl.lowrank_setup_times()
            res["lowrank"] = {"gram_product_ms_each": ph["gram_products_ms"] / 2.0, "orthonormalisation_ms": ph["orthonormalisation_ms"],
                              "sigma2_1": hl.nonrigid_step()}
            return res
        out[name] = guarded(run)
    return out


def parity_probes():
    from probreg_b200 import bcpd, cpd

    out = {}
    src, tgt = deformed(2000)

    def lowrank_vs_dense():
        a = cpd.NonRigidCPD(src, beta=2.0, lmd=2.0)
        ra = a.registration(tgt, w=0.05, maxiter=6, tol=-1.0)
        b = cpd.NonRigidCPD(src, beta=2.0, lmd=2.0, low_rank=100)
        rb = b.registration(tgt, w=0.05, maxiter=6, tol=-1.0)
        q = rb.transformation.q
        return {"sigma2_dense": ra.sigma2, "sigma2_lowrank": rb.sigma2, "rel": rb.sigma2 / ra.sigma2 - 1.0,
                "max_moved_diff": float(np.abs(a.moved_source() - b.moved_source()).max()),
                "orthonormality_defect": float(np.abs(q.T.dot(q) - np.identity(q.shape[1])).max()),
                "expect": "CPU emulation of the same code: sigma2_dense 0.009435964, rel 4.1e-7, max_moved_diff 2.5e-5; bars: rel 1e-5, moved 1e-4"}

    def bcpd_vs_fixture():
        g = np.load(os.path.join(ROOT, "tests", "golden", "bcpd.npz"))
        res = {}
        for tag in ("b", "c"):
            x = g[str(g[tag + "_target"])]
            es = bcpd.CombinedBCPD(g["source"]).expectation_step(g["t_source"], x, float(g[tag + "_scale"]), g[tag + "_alpha"],
                                                                 g[tag + "_sdiag"], float(g[tag + "_sigma2"]), float(g[tag + "_w"]))
            ref = g[tag + "_nu"]
            ok = ref > 1e-9
            res[tag] = {"max_rel_err_nu": float(np.abs(es.nu[ok] / ref[ok] - 1.0).max()),
                        "dead_columns_match": bool(np.array_equal(es.nu_d == 0, g[tag + "_nu_d"] == 0))}
        res["expect"] = "max_rel_err_nu < 5e-5, dead columns match (tests/test_zz_bcpd.py)"
        return res

    def constrained():
        idx = np.arange(0, 2000, 20)
        r = cpd.registration_cpd(src, tgt, "nonrigid_constrained", maxiter=4, tol=-1.0, beta=1.0, lmd=1.5, alpha=1e-2,
                                 idx_source=idx, idx_target=idx)
        return {"sigma2": r.sigma2, "expect": "CPU emulation of the same code: 0.0110745100 (reference arithmetic, oracle: within 1e-5)"}

    out["lowrank_vs_dense_2000"] = guarded(lowrank_vs_dense)
    out["bcpd_estep_vs_reference_fixture"] = guarded(bcpd_vs_fixture)
    out["constrained_device_priors_2000"] = guarded(constrained)
    return out


def main():
    t0 = time.perf_counter()
    res = {"affine_250k": guarded(affine_250k), "lowrank_50k": guarded(lowrank_50k), "bcpd_estep_100k": guarded(bcpd_estep_100k)}
    res["first_hardware_run_probes"] = guarded(parity_probes)
    res["variants"] = guarded(variants)
    res["seconds"] = time.perf_counter() - t0
    print(json.dumps(res))


if __name__ == "__main__":
    main()
