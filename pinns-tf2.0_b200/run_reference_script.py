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
    if os.path.isdir(os.path.join(ref_root, "PINNs")):        # IRK tables of the upstream repository (discrete-time models)
        os.symlink(os.path.join(ref_root, "PINNs"), os.path.join(wd, "PINNs"))
    for eqn in ("1d-burgers", "1dcomplex-schrodinger"):
        d = os.path.join(wd, eqn)
        os.makedirs(d)
        for f in os.listdir(os.path.join(HERE, eqn)):
            os.symlink(os.path.join(HERE, eqn, f), os.path.join(d, f))
        if os.path.isdir(os.path.join(ref_root, eqn, "data")):
            os.symlink(os.path.join(ref_root, eqn, "data"), os.path.join(d, "data"))
    return wd


def normalise_indentation(src):
    """Rebuild a consistent 4-space block structure for a script whose indentation is broken (the shipped
    1d-burgers/ide_cont_burgers.py mixes 2- and 4-space levels and U+00A0).  Blocks are reconstructed from the
    statement structure, not from the raw widths:
      * a new block opens only after a line that ends with ':';
      * inside a class, every `def ...(self...)` is a method of that class, whatever its raw indentation;
      * a statement leaves a nested block only if its raw indentation is smaller than that block's AND not larger than the
        enclosing block's; it never leaves a method body or (while indented) a top-level compound statement;
      * continuation lines (open brackets / trailing backslash) stay attached to their statement.
    Statements, names and expressions are untouched."""
    lines = src.replace("\u00a0", " ").replace("\t", "    ").split("\n")
    out = []
    stack = [(0, 0)]            # (raw width of the block's first statement, logical level)
    in_class = False
    prev_opens = False          # previous code line ended with ':'
    prev_was_def = False
    depth = 0                   # bracket depth for continuation lines
    backslash = False
    level = 0
    for line in lines:
        stripped = line.strip()
        if depth > 0 or backslash:                       # continuation of the previous statement
            out.append(" " * (4 * level + 8) + stripped)
            code = stripped.split("#")[0]
            depth += sum(code.count(c) for c in "([{") - sum(code.count(c) for c in ")]}")
            backslash = stripped.endswith("\\")
            if depth <= 0 and not backslash:
                depth = 0
                prev_opens = code.rstrip().endswith(":")
            continue
        if not stripped or stripped.startswith("#"):
            out.append((" " * (4 * level) + stripped) if stripped else "")
            continue
        raw = len(line) - len(line.lstrip(" "))
        is_method = in_class and stripped.startswith("def ") and "(self" in stripped
        if stripped.startswith("class ") and raw == 0:
            in_class, level, stack = True, 0, [(0, 0)]
        elif is_method:
            level, stack = 1, [(0, 0), (raw, 1)]
        elif in_class and raw == 0 and not prev_opens:
            in_class, level, stack = False, 0, [(0, 0)]   # top-level code resumes after the class
        elif prev_opens:
            level = stack[-1][1] + 1
            stack.append((raw, level))
        else:
            floor = 3 if in_class else 1                  # never leave a method body / keep indented code in its block
            while len(stack) > floor and raw < stack[-1][0] and raw <= stack[-2][0]:
                stack.pop()
            if not in_class and raw == 0:
                stack = [(0, 0)]
            level = stack[-1][1]
        if is_method or (stripped.startswith("class ") and raw == 0):
            pass
        out.append(" " * (4 * level) + stripped)
        code = stripped.split("#")[0]
        depth = sum(code.count(c) for c in "([{") - sum(code.count(c) for c in ")]}")
        depth = max(depth, 0)
        backslash = stripped.endswith("\\")
        prev_opens = depth == 0 and not backslash and code.rstrip().endswith(":")
        if is_method and prev_opens:
            # the method body is one level below the def; its first statement defines the body's raw width
            stack = [(0, 0), (raw, 1)]
    return "\n".join(out)


def main(argv):
    if len(argv) < 2:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[1])
    wd = prepare_workdir(script)
    for pth in (os.path.join(HERE, "shims"), os.path.join(wd, os.path.basename(os.path.dirname(script)))):
        sys.path.insert(0, pth)
    # result artefacts (hp.json, fields.npz, figure) land under <launch directory>/<eqn>/results/, not in the scratch directory
    os.environ.setdefault("PINN_RESULTS_ROOT", os.getcwd())
    # optional script arguments that name files (the hp.json of `script.py hp.json`) were given relative to the caller's
    # directory, not to the scratch directory the script runs in
    rest = [os.path.abspath(a) if os.path.exists(a) else a for a in argv[2:]]
    os.chdir(wd)
    sys.argv = [script] + rest
    src = open(script, encoding="utf-8").read()
    try:
        compile(src, script, "exec")
    except SyntaxError:
        # the shipped file does not parse: repair the indentation in memory, change nothing else
        src = normalise_indentation(src)
        fixed = os.path.join(wd, os.path.basename(script))
        with open(fixed, "w", encoding="utf-8") as f:
            f.write(src)
        script = fixed
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
