"""Multi-GPU path on real devices: torchrun, one rank per GPU, NCCL all-reduce inside libcpd_b200.so (skipped when fewer GPUs
are visible than ranks), and the NCCL-free part of it -- sharding + the fused peer-memory exchange -- with two ranks on one GPU."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from probreg_b200 import _cabi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_registration_matches_oracle(world):
    if _cabi.lib().cpd_device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DIST_OK world=%d" % world in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


@pytest.mark.timeout(600)
def test_two_ranks_share_one_gpu():
    """The sharded EM loop with the fused peer-memory exchange, two ranks on ONE device (runs on the single-GPU round-end box,
    where the NCCL tests above skip): see tests/dist_worker_1gpu.py."""
    env = dict(os.environ, CPD_B200_P2P_TIMEOUT_S="20")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29521", os.path.join(ROOT, "tests", "dist_worker_1gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=500, cwd=ROOT, env=env)
    assert r.returncode == 0 and "DIST1GPU_OK world=2" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
