#!/usr/bin/env python
"""Target-sharded rigid CPD over all GPUs of a node:
   python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/multi_gpu.py [points]"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probreg_b200 import cpd  # noqa: E402
from probreg_b200 import dist as pdist  # noqa: E402
from probreg_b200.synthetic import synthetic_pair  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = pdist.Communicator.from_torch(local)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
source, target = synthetic_pair(n)                      # every rank holds the full clouds; targets are sharded inside
t0 = time.time()
res = cpd.registration_cpd(source, target, maxiter=20, tol=-1.0, comm=comm)
if comm.rank == 0:
    print("%d points on %d GPUs: 20 EM iterations in %.2f s, sigma2 %.4e" % (n, comm.world_size, time.time() - t0, res.sigma2))
dist.destroy_process_group()
