"""pytest plugin (``-p oracle_backend_plugin`` with tests/ on PYTHONPATH), TEST TOOLING ONLY: replaces ``pinn_cabi.Pinn`` by the
oracle-backed stand-in so that the PYTHON of the ``-m gpu`` test files can be executed on a box without a GPU.  It proves
nothing about the CUDA path; it catches typos / wrong keys / API misuse in GPU test code before GPU time is spent on them
(tests/test_gpu_testcode_dryrun.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "pinns-tf2.0_b200", "utils")]
import oracle_backend  # noqa: E402
import pinn_cabi  # noqa: E402

pinn_cabi.Pinn = oracle_backend.OraclePinn
pinn_cabi.load = lambda: None
