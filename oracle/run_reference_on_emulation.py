"""Run an UNMODIFIED reference script end to end on the CPU, with TensorFlow replaced by oracle/tf_emulation.

    python oracle/run_reference_on_emulation.py /root/reference/1d-burgers/inf_cont_burgers.py [hp.json]

TEST INFRASTRUCTURE (see oracle/__init__.py): this is how "what does the reference print / how long does the reference take on
this host" is answered in a container where tensorflow==2.0.0-rc0 cannot be installed.  Everything the script imports is the
reference's own file (utils/*.py, <eqn>/*util.py, data, IRK tables) except
  * tensorflow / tensorflow_probability -> oracle/tf_emulation (torch CPU fp64 behind TF-2.0 tape/Keras/Adam semantics),
  * pyDOE.lhs                            -> the classic-LHS stand-in of pinns-tf2.0_b200/shims (pyDOE is not installed),
  * matplotlib / mpl_toolkits           -> inert stubs (figures are cosmetic; hp.json is still written by saveResultDir).
The script runs from a scratch directory laid out like the reference checkout (read-only checkout, results go to the scratch
directory).  Scripts that do not parse as shipped (ide_cont_burgers.py) are re-indented in memory exactly like
pinns-tf2.0_b200/run_reference_script.py does for the B200 run.
"""
import importlib.util
import os
import runpy
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class _Inert(types.ModuleType):
    """Stands in for matplotlib objects: any attribute, call, index or iteration yields another inert object."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return _Inert(self.__name__ + "()")

    def __getitem__(self, k):
        return _Inert(self.__name__ + "[]")

    def __iter__(self):
        return iter((_Inert(self.__name__ + "[0]"), _Inert(self.__name__ + "[1]")))

    def update(self, *a, **k):
        return None


def prepare_workdir(script):
    ref_root = os.path.dirname(os.path.dirname(os.path.abspath(script)))
    wd = tempfile.mkdtemp(prefix="pinn_ref_emul_")
    for name in ("utils", "PINNs"):
        if os.path.isdir(os.path.join(ref_root, name)):
            os.symlink(os.path.join(ref_root, name), os.path.join(wd, name))
    for eqn in ("1d-burgers", "1dcomplex-schrodinger"):
        d = os.path.join(wd, eqn)
        os.makedirs(os.path.join(d, "results"))
        for f in os.listdir(os.path.join(ref_root, eqn)):
            if f != "results":
                os.symlink(os.path.join(ref_root, eqn, f), os.path.join(d, f))
    return wd


def main(argv):
    if len(argv) < 2:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[1])
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "mpl_toolkits.axes_grid1"):
        sys.modules[name] = _Inert(name)
    sys.path.insert(0, os.path.join(HERE, "tf_emulation"))
    # only pyDOE from the shims directory (its `tensorflow` shim must NOT shadow the emulation)
    spec = importlib.util.spec_from_file_location("pyDOE", os.path.join(ROOT, "pinns-tf2.0_b200", "shims", "pyDOE", "__init__.py"))
    pydoe = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pydoe)
    sys.modules["pyDOE"] = pydoe
    wd = prepare_workdir(script)
    os.chdir(wd)
    # `python <eqn>/<script>.py` puts the script's directory first on sys.path (inf_cont_schrodinger.py relies on it)
    sys.path.insert(0, os.path.join(wd, os.path.basename(os.path.dirname(script))))
    sys.argv = [script] + argv[2:]
    src = open(script, encoding="utf-8").read()
    try:
        compile(src, script, "exec")
    except SyntaxError:
        spec = importlib.util.spec_from_file_location("_runner", os.path.join(ROOT, "pinns-tf2.0_b200", "run_reference_script.py"))
        runner = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(runner)
        script = os.path.join(wd, os.path.basename(script))
        with open(script, "w", encoding="utf-8") as f:
            f.write(runner.normalise_indentation(src))
    t0 = time.time()
    runpy.run_path(script, run_name="__main__")
    print("[reference on emulated TF] %s finished in %.1f s; scratch directory %s" % (os.path.basename(script), time.time() - t0, wd))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
