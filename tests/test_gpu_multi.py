"""GPU (>= 2 devices): sharded evaluation over the fused NVLink push exchange and over NCCL against the single-GPU evaluation of
the whole point set (tests/mgpu_check.py under torch.distributed.run).  Skipped on a one-GPU box; bench.py's `parity_check`
block covers the same ground inside every multi-GPU bench run."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs on the box")
def test_sharded_matches_single_gpu_on_two_ranks():
    port = 29700 + os.getpid() % 200
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert r.stdout.count("mgpu_check ok") == 4, r.stdout[-2000:]
