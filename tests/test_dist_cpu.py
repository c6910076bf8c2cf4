"""Host-side logic of the multi-GPU path on CPU: two gloo ranks (no GPU needed).

The device side of target sharding is "each rank runs the same E-step on its shard with the
GLOBAL N in the outlier constant, then sums".  Here the oracle plays the device: it checks the
sharding arithmetic (shard bounds, global N, frame origin, what is summed and what is
concatenated) and the rendezvous that hands the NCCL unique id from rank 0 to the others.
"""
import os
import socket

import numpy as np
import pytest

from oracle import cpd_oracle as orc
from probreg_b200 import dist as pdist


def test_shard_bounds_cover_and_balance():
    for n in (1, 7, 1000, 100003):
        for w in (1, 2, 3, 8):
            edges = [pdist.shard_bounds(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges[:-1], edges[1:]):
                assert b == c
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as tdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch

        comm = pdist.Communicator.from_torch(device=0)
        assert (comm.rank, comm.world_size) == (rank, world)
        uid = comm.unique_id()                      # created on rank 0 by libcpd_b200, broadcast
        assert isinstance(uid, bytes) and len(uid) == 128
        gathered = [None] * world
        tdist.all_gather_object(gathered, uid)
        assert all(g == gathered[0] for g in gathered)
        uid2 = comm.unique_id()
        assert uid2 != uid                          # every call mints a fresh id

        src, tgt = orc.synthetic_pair(600)
        outl = (np.random.default_rng(3).random((50, 3)) - 0.5) * 3 + tgt.mean(0)
        tgt = np.r_[tgt, outl]
        lo, hi = comm.shard_bounds(tgt.shape[0])
        origin = comm.frame_origin(tgt)
        assert np.array_equal(origin, tgt.mean(0))
        # one EM iteration, sharded: local E-step with the global N, all-reduce, replicated M-step
        s2, w = 0.01, 0.2
        es = orc.expectation_step(src, tgt[lo:hi], s2, w, n_global=tgt.shape[0])
        buf = torch.from_numpy(np.r_[es.p1, es.px.ravel()].copy())
        tdist.all_reduce(buf)
        p1 = buf[: src.shape[0]].numpy()
        px = buf[src.shape[0]:].numpy().reshape(-1, 3)
        pt1_parts = [None] * world
        tdist.all_gather_object(pt1_parts, es.pt1)
        full = orc.expectation_step(src, tgt, s2, w)
        np.testing.assert_allclose(p1, full.p1, rtol=1e-12)
        np.testing.assert_allclose(px, full.px, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(np.concatenate(pt1_parts), full.pt1, rtol=1e-14)
        a = orc.mstep_rigid(src, tgt, orc.Estep(np.concatenate(pt1_parts), p1, px, float(p1.sum())))
        b = orc.mstep_rigid(src, tgt, full)
        np.testing.assert_allclose(a.params[0], b.params[0], atol=1e-12)
        assert a.sigma2 == pytest.approx(b.sigma2, rel=1e-11)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        tdist.destroy_process_group()


def test_two_rank_gloo_sharding_model():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(30)
    assert res == [(0, "ok"), (1, "ok")], res
