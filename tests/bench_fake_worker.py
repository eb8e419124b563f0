"""Worker for test_bench_dryrun.py::test_bench_two_ranks_gloo: runs bench.main() on one rank of a 2-rank gloo job with the
fake backend (no GPU).  Launched as a plain subprocess with RANK/WORLD_SIZE/MASTER_* in the environment, exactly what
torch.distributed.run would set."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, HERE, os.path.join(ROOT, "pinns-tf2.0_b200", "utils")]

import torch            # noqa: E402
import pinn_cabi        # noqa: E402
import sharding         # noqa: E402
import bench            # noqa: E402
from test_bench_dryrun import FakePinn  # noqa: E402

pinn_cabi.Pinn = FakePinn
pinn_cabi.host_alloc = lambda n: (np.zeros(n), C.c_void_p(0))
pinn_cabi.nccl_unique_id = lambda: b"N" * 128
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.set_device = lambda *a, **k: None
sharding.connect_p2p = lambda dist, p, world: False
sys.argv = ["bench.py", "--gpus", "2", "--steps", "3", "--warmup", "3"] + sys.argv[1:]
bench.main()
