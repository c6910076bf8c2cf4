"""Multi-GPU path on real devices: torchrun, one rank per GPU, NCCL all-reduce inside libcpd_b200.so.
Skipped when fewer than two GPUs are visible (the single-GPU round-end box)."""
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
