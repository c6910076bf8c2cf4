"""Worker for tests/test_dist_gpu.py: run under torchrun with one rank per GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def say(msg):
    sys.stderr.write("[rank %s] %s\n" % (os.environ.get("RANK", "?"), msg))
    sys.stderr.flush()


def main():
    import torch
    import torch.distributed as tdist

    from oracle import cpd_oracle as orc
    from probreg_b200 import cpd
    from probreg_b200 import dist as pdist

    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = pdist.Communicator.from_torch(local)
    rank, world = comm.rank, comm.world_size

    src, tgt = orc.synthetic_pair(3001)
    outl = (np.random.default_rng(3).random((211, 3)) - 0.5) * 3 + tgt.mean(0)
    tgt = np.ascontiguousarray(np.r_[tgt, outl])
    say("init done")
    # (1) sharded E-step: p1/px/n_p are global sums on every rank, pt1 is the local shard
    r = cpd.RigidCPD(src, comm=comm)
    es = r.expectation_step(src, tgt, 0.01, 0.15)
    say("estep done")
    ref = orc.expectation_step(src, tgt, 0.01, 0.15)
    lo, hi = comm.shard_bounds(tgt.shape[0])
    np.testing.assert_allclose(es.pt1, ref.pt1[lo:hi], rtol=2e-5, atol=1e-12)
    np.testing.assert_allclose(es.p1, ref.p1, rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(es.px, ref.px, rtol=2e-5, atol=2e-5 * np.abs(ref.px).max())
    assert abs(es.n_p - ref.n_p) < 1e-6 * ref.n_p
    # (2) sharded registration == oracle, and identical on every rank
    for kind, kw in (("rigid", {}), ("affine", {}), ("rigid", {"update_scale": False})):
        res = cpd.registration_cpd(src, tgt, kind, w=0.1, maxiter=12, tol=-1.0, comm=comm, **kw)
        say("registration %s done" % kind)
        oref, _ = orc.registration(src, tgt, kind, w=0.1, maxiter=12, tol=-1.0, **kw)
        lin = res.transformation.rot if kind == "rigid" else res.transformation.b
        np.testing.assert_allclose(lin, oref.params[0], atol=1e-5)
        np.testing.assert_allclose(res.transformation.t, oref.params[1], atol=1e-5)
        assert abs(res.sigma2 - oref.sigma2) <= 1e-6 * oref.sigma2, (res.sigma2, oref.sigma2)
        box = [None] * world
        tdist.all_gather_object(box, (float(res.sigma2), lin.tolist()))
        assert all(b == box[0] for b in box), "ranks disagree"
    # (3) stand-alone M-step from a sharded EstepResult
    say("mstep start")
    m = r.maximization_step(tgt, es)
    say("mstep done")
    mref = orc.mstep_rigid(src, tgt, ref)
    np.testing.assert_allclose(m.transformation.rot, mref.params[0], atol=1e-5)
    # (4) low-rank non-rigid: the G X products are sharded over rows (integer digits on the tensor cores: the factors do not depend
    #     on the row tiling), the K x K M-step is replicated on the all-reduced p1 / px
    f = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.7], [0.3, 0.0, 1.0]])
    bent = np.ascontiguousarray(src + 0.03 * np.sin(2 * np.pi * src.dot(f)))
    one = cpd.NonRigidCPD(src, beta=1.5, lmd=2.0, low_rank=60, device=local)
    r1 = one.registration(bent, w=0.05, maxiter=5, tol=-1.0)
    many = cpd.NonRigidCPD(src, beta=1.5, lmd=2.0, low_rank=60, comm=comm)
    rn = many.registration(bent, w=0.05, maxiter=5, tol=-1.0)
    say("low-rank non-rigid done")
    assert np.array_equal(rn.transformation.q, r1.transformation.q) and np.array_equal(rn.transformation.bcore, r1.transformation.bcore)
    assert abs(rn.sigma2 - r1.sigma2) <= 1e-6 * r1.sigma2, (rn.sigma2, r1.sigma2)
    np.testing.assert_allclose(many.moved_source(), one.moved_source(), atol=1e-6)
    box = [None] * world
    tdist.all_gather_object(box, (float(rn.sigma2), float(np.abs(many.moved_source()).sum())))
    assert all(b == box[0] for b in box), "ranks disagree (low-rank)"
    tdist.barrier()
    if rank == 0:
        print("DIST_OK world=%d" % world)
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
