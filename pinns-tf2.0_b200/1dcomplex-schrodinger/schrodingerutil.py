"""Data preparation for the NLS problem (restates 1dcomplex-schrodinger/schrodingerutil.py:21-61 of the reference) and the
run's artefact writer (shares the helpers of 1d-burgers/burgersutil.py)."""
import importlib.util
import os

import sys

import numpy as np
import scipy.io

sys.path.append("utils")   # the reference modules do the same (burgersutil.py:24, schrodingerutil.py:18); scripts rely on it

_b = importlib.util.spec_from_file_location("_burgersutil_lhs", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                                                             "1d-burgers", "burgersutil.py"))
_m = importlib.util.module_from_spec(_b); _b.loader.exec_module(_m)
_lhs = _m._lhs


def prep_data(path, N_0, N_b, N_f, noise):
    data = scipy.io.loadmat(path)
    t = data["tt"].flatten()[:, None]
    x = data["x"].flatten()[:, None]
    Exact = data["uu"]
    Exact_u, Exact_v = np.real(Exact), np.imag(Exact)
    Exact_h = np.sqrt(Exact_u ** 2 + Exact_v ** 2)
    X, T = np.meshgrid(x, t)
    X_star = np.hstack((X.flatten()[:, None], T.flatten()[:, None]))
    u_star = Exact_u.T.flatten()[:, None]
    v_star = Exact_v.T.flatten()[:, None]
    h_star = Exact_h.T.flatten()[:, None]
    lb = np.array([-5.0, 0.0])
    ub = np.array([5.0, np.pi / 2])
    idx_x = np.random.choice(x.shape[0], N_0, replace=False)
    x0 = x[idx_x, :]
    u0, v0 = Exact_u[idx_x, 0:1], Exact_v[idx_x, 0:1]
    idx_t = np.random.choice(t.shape[0], N_b, replace=False)
    tb = t[idx_t, :]
    X0 = np.concatenate((x0, 0 * x0), 1)
    H0 = np.hstack((u0, v0))
    X_f = lb + (ub - lb) * _lhs(2, N_f)
    return x, t, X, T, Exact_u, Exact_v, Exact_h, X_star, u_star, v_star, h_star, X_f, ub, lb, tb, x0, u0, v0, X0, H0


def plot_inf_cont_results(X_star, u_pred, v_pred, h_pred, Exact_h, X, T, x, t, ub, lb, x0, tb, save_path=None, save_hp=None):
    """Artefacts of the Schrodinger run (reference figure: schrodingerutil.py:63-140): |h| on the (t, x) grid, the three time
    slices t[75], t[100], t[125], the training point locations; drawn only when matplotlib exists."""
    U, V, H = (_m._grid(X_star, a, X, T) for a in (u_pred, v_pred, h_pred))
    rows = (75, 100, 125)
    X_train = np.vstack([np.concatenate((x0, 0 * x0), 1), np.concatenate((0 * tb + lb[0], tb), 1),
                         np.concatenate((0 * tb + ub[0], tb), 1)])
    return _m._finish(save_path, save_hp,
                      lambda pl: _m._slices(pl, x, [Exact_h[:, r] for r in rows], [H[r] for r in rows], ["t = %.2f" % t[r, 0] for r in rows]),
                      U_pred=U, V_pred=V, H_pred=H, Exact_h=Exact_h, x=x, t=t, X_train=X_train, slice_rows=np.array(rows),
                      rel_l2_error_h=np.linalg.norm(Exact_h.T - H) / np.linalg.norm(Exact_h))
