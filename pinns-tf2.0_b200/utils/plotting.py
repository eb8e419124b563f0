"""Result artefacts of a run: ``<save_path>/results/<YYYYmmdd-HHMMSS>-<script>/`` holding ``hp.json`` (the
hyper-parameters, as the reference writes them, utils/plotting.py:8-16), ``fields.npz`` (every array the figure would show, so
the figure can be redrawn anywhere) and, when matplotlib is importable, ``graph.png`` / ``graph.pdf``.

The reference draws through LaTeX/pgf; neither LaTeX nor matplotlib is part of the training hot path and neither exists on the
GPU image, so drawing is optional here and never an error.  ``newfig`` / ``savefig`` / ``figsize`` keep the reference names for
user code that imports them."""
import json
import os
import sys
from datetime import datetime

import numpy as np

try:                                        # optional; plain Agg backend, no LaTeX
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
except Exception:                           # pragma: no cover - depends on the image
    plt = None

_pending = {}                               # arrays registered by the plot_* helpers for the next saveResultDir


def have_matplotlib():
    return plt is not None


def figsize(scale, nplots=1):
    width = 390.0 / 72.27 * scale           # LaTeX text width of the upstream paper, in inches
    return [width, nplots * width * (np.sqrt(5.0) - 1.0) / 2.0]


def newfig(width, nplots=1):
    if plt is None:
        return None, None
    fig = plt.figure(figsize=figsize(width, nplots))
    return fig, fig.add_subplot(111)


def savefig(filename, crop=True):
    if plt is None:
        return []
    kw = {"bbox_inches": "tight", "pad_inches": 0} if crop else {}
    written = []
    for ext in ("pdf", "png"):
        plt.savefig("%s.%s" % (filename, ext), **kw)
        written.append("%s.%s" % (filename, ext))
    return written


def stage_fields(**arrays):
    """Remember the arrays behind the current figure; the next saveResultDir writes them to fields.npz."""
    _pending.clear()
    _pending.update({k: np.asarray(v) for k, v in arrays.items() if v is not None})


def _jsonable(v):
    if isinstance(v, (np.floating, np.integer)):
        return v.item()
    if isinstance(v, np.ndarray):
        return v.tolist()
    raise TypeError("hp value of type %s is not JSON serialisable" % type(v).__name__)


def saveResultDir(save_path, save_hp):
    if int(os.environ.get("RANK", "0")) != 0:           # one artefact directory per job, not per rank
        return None
    script = os.path.splitext(os.path.basename(sys.argv[0]))[0]
    if not os.path.isabs(save_path):                    # run_reference_script.py runs the script from a scratch directory
        save_path = os.path.join(os.environ.get("PINN_RESULTS_ROOT", "."), save_path)
    res_dir = os.path.join(save_path, "results", "%s-%s" % (datetime.now().strftime("%Y%m%d-%H%M%S"), script))
    os.makedirs(res_dir, exist_ok=True)
    print("Saving results to directory ", os.path.realpath(res_dir))
    savefig(os.path.join(res_dir, "graph"))
    with open(os.path.join(res_dir, "hp.json"), "w") as f:
        json.dump(save_hp, f, default=_jsonable)
    if _pending:
        np.savez_compressed(os.path.join(res_dir, "fields.npz"), **_pending)
        _pending.clear()
    return res_dir
