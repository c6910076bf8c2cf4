"""Worker for test_two_ranks_share_one_gpu (tests/test_dist_gpu.py): TWO ranks on ONE GPU.

The round-end GPU box has a single device, where NCCL cannot run two ranks (it refuses duplicate devices) -- but the part of the
multi-rank path this library wrote itself does not need NCCL: targets sharded over ranks, sources replicated, the 32 moment sums
exchanged inside moments_p2p_kernel through peer-mapped mailboxes (cudaIpc works between two processes on one device too).
torch.distributed/gloo is the rendezvous only.  What is checked is what the multi-GPU run relies on: sharded EM iterations equal
the single-rank oracle and are bit-identical on both ranks.  Launch: torchrun --nproc-per-node 2 tests/dist_worker_1gpu.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as tdist

    from oracle import cpd_oracle as orc
    from probreg_b200 import _cabi
    from probreg_b200 import dist as pdist

    tdist.init_process_group("gloo")
    rank, world = tdist.get_rank(), tdist.get_world_size()

    def gather(obj):
        box = [None] * world
        tdist.all_gather_object(box, obj)
        return box

    src, tgt = orc.synthetic_pair(2501)
    outl = (np.random.default_rng(3).random((173, 3)) - 0.5) * 3 + tgt.mean(0)
    tgt = np.ascontiguousarray(np.r_[tgt, outl])
    n = tgt.shape[0]
    lo, hi = pdist.shard_bounds(n, rank, world)
    origin = pdist.Communicator(rank, world).frame_origin(tgt)
    for kind, tf_kind in (("rigid", _cabi.TF_RIGID), ("affine", _cabi.TF_AFFINE)):
        h = _cabi.Handle(3, device=0)
        h.set_source(src)
        h.set_target(tgt[lo:hi], n_global=n, frame_origin=origin)
        h.p2p_attach(gather(h.p2p_local_handle()), world, rank)
        gather(0)                                            # nobody steps before everybody has mapped everybody
        s2 = float(orc.sigma2_init(src, tgt))
        h.set_state(tf_kind, True, 0.1, np.identity(3), np.zeros(3), 1.0, s2, 1.0 + n * 1.5 * np.log(s2))
        out = None
        for _ in range(10):
            out = h.em_step()
        oref, _ = orc.registration(src, tgt, kind, w=0.1, maxiter=10, tol=-1.0)
        lin = out[0] if kind == "affine" else out[0]
        np.testing.assert_allclose(lin, oref.params[0], atol=1e-5)
        np.testing.assert_allclose(out[1], oref.params[1], atol=1e-5)
        assert abs(out[3] - oref.sigma2) <= 1e-6 * oref.sigma2, (out[3], oref.sigma2)
        box = gather((float(out[3]), np.asarray(out[0]).tolist(), np.asarray(out[1]).tolist()))
        assert all(b == box[0] for b in box), "ranks disagree"
        gather(0)
        h.p2p_detach()
        h.close()
    if rank == 0:
        print("DIST1GPU_OK world=%d" % world)
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
