"""CPU: dry run of the GPU parity tests' own Python against the oracle stand-in (tests/oracle_backend_plugin.py).  Only tests
whose assertions the stand-in can serve are selected (no timing, zero-copy, device-tanh, error-path or empty-set cases)."""
import os
import subprocess
import sys

import torch
import pytest

from conftest import ROOT

SELECTION = [
    ("tests/test_gpu_burgers.py", "golden or probes or adam_trajectory or identification"),
    ("tests/test_gpu_nls.py", "both_ic or probes or adam_trajectory"),
    ("tests/test_gpu_disc.py", "golden or synthetic or surface"),
    ("tests/test_gpu_surface.py", "fit_matches or identification_surface or schrodinger_surface"),
]


@pytest.mark.skipif(torch.cuda.is_available(), reason="with a GPU the real tests run instead")
@pytest.mark.parametrize("path,expr", SELECTION)
def test_gpu_test_code_runs_on_the_oracle_stand_in(path, expr):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "pytest", "-p", "oracle_backend_plugin", os.path.join(ROOT, path), "-m", "gpu", "-q",
                        "-x", "-k", expr, "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    tail = r.stdout[-2500:] + r.stderr[-1500:]
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, tail
