"""CPU: the shipped path must not depend on the checker.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may touch oracle/; nothing under the package may, and the native sources must not carry a host
fallback for the loss/gradient arithmetic."""
import ast
import os
import re
import subprocess
import sys

from conftest import ROOT

PKG = os.path.join(ROOT, "pinns-tf2.0_b200")


def _py_files(top):
    for d, _, fs in os.walk(top):
        if "__pycache__" in d:
            continue
        for f in fs:
            if f.endswith(".py"):
                yield os.path.join(d, f)


def _imports(path):
    tree = ast.parse(open(path, encoding="utf-8").read(), path)
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name
        elif isinstance(node, ast.ImportFrom):
            yield node.module or ""


def test_package_never_imports_the_oracle_or_torch_autograd():
    for path in _py_files(PKG):
        mods = list(_imports(path))
        assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), path
        src = open(path, encoding="utf-8").read()
        assert "reference_port" not in src and "taylor" not in src.lower().replace("taylor-mode", ""), path
        # torch is plumbing only (torch.distributed in sharding.py); no autograd / nn in the product
        assert not re.search(r"torch\.(autograd|nn)\b|\.backward\(", src), path


def test_native_sources_have_no_host_arithmetic_fallback():
    csrc = os.path.join(PKG, "csrc")
    for f in os.listdir(csrc):
        src = open(os.path.join(csrc, f), encoding="utf-8").read()
        assert "oracle" not in src.lower(), f
        # tanh evaluated on the host would be the signature of a CPU restatement of the network inside the library
        host_code = re.sub(r"__global__[\s\S]*?\n}\n", "", src) if f.endswith(".cu") else ""
        assert not re.search(r"\bstd::tanh\b", host_code), f


def test_bench_and_smoke_use_the_oracle_only_as_checker():
    bench = open(os.path.join(ROOT, "bench.py"), encoding="utf-8").read()
    # the only oracle entry points in bench.py are the CPU timing helpers (one contiguous section: cpu_threads ... _port_baseline)
    # shared by cpu_baseline, the extras' cpu_baselines and --impl reference
    users = [m.start() for m in re.finditer(r"from oracle|import oracle", bench)]
    assert users, "bench.py must time the oracle for cpu_baseline"
    lo, hi = bench.index("def cpu_threads"), bench.index("def run_reference_arm")
    body = bench[lo:hi]
    assert all(lo < u < hi for u in users), "oracle imported outside the CPU timing section"
    assert "pinn_cabi" not in body                      # the CPU leg never touches the product, and vice versa
    entry = open(os.path.join(ROOT, "__graft_entry__.py"), encoding="utf-8").read()
    build_body = entry[entry.index("def build"):entry.index("def smoke")]
    assert not re.search(r"from oracle|import oracle", build_body)


def test_importing_the_mirror_does_not_load_the_oracle():
    code = ("import sys; sys.path[:0]=[%r, %r]; import neuralnetwork, custom_lbfgs, logger, pinn_cabi, sharding, plotting; "
            "bad=[m for m in sys.modules if m=='oracle' or m.startswith('oracle.')]; print(bad); sys.exit(1 if bad else 0)"
            % (os.path.join(PKG, "utils"), os.path.join(PKG, "shims")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
