"""Data preparation for the continuous-time Burgers problems (restates 1d-burgers/burgersutil.py:27-36, 63-75, 99-131
of the reference; the discrete-time/IRK branches are out of scope).  Plot helpers degrade to no-ops when
matplotlib is missing (cosmetic)."""
import sys

import numpy as np
import scipy.io

sys.path.append("utils")   # the reference modules do the same (burgersutil.py:24, schrodingerutil.py:18); scripts rely on it

try:
    from pyDOE import lhs
except Exception:  # pragma: no cover
    lhs = None


def _lhs(d, n):
    if lhs is not None:
        return lhs(d, n)
    cut = np.linspace(0, 1, n + 1)
    u = np.random.rand(n, d)
    rd = u * (cut[1:] - cut[:n])[:, None] + cut[:n][:, None]
    out = np.zeros_like(rd)
    for j in range(d):
        out[:, j] = rd[np.random.permutation(range(n)), j]
    return out


def prep_data(path, N_u=None, N_f=None, N_n=None, q=None, ub=None, lb=None, noise=0.0, idx_t_0=None, idx_t_1=None,
              N_0=None, N_1=None):
    if N_n is not None or N_0 is not None:
        raise NotImplementedError("discrete-time (IRK) data preparation is out of scope of the hot path (SURVEY 8(f)2)")
    data = scipy.io.loadmat(path)
    t = data["t"].flatten()[:, None]
    x = data["x"].flatten()[:, None]
    Exact_u = np.real(data["usol"]).T
    X, T = np.meshgrid(x, t)
    X_star = np.hstack((X.flatten()[:, None], T.flatten()[:, None]))
    u_star = Exact_u.flatten()[:, None]
    # identification: N_u random grid points, `noise` ignored exactly like the reference (burgersutil.py:72-75)
    idx = np.random.choice(X_star.shape[0], N_u, replace=False)
    X_u_train, u_train = X_star[idx, :], u_star[idx, :]
    lb, ub = X_star.min(axis=0), X_star.max(axis=0)
    if N_f is None:
        return x, t, X, T, Exact_u, X_star, u_star, X_u_train, u_train, ub, lb
    # inference: initial + boundary points, then a Latin hypercube of collocation points (burgersutil.py:104-129)
    xx1 = np.hstack((X[0:1, :].T, T[0:1, :].T)); uu1 = Exact_u[0:1, :].T
    xx2 = np.hstack((X[:, 0:1], T[:, 0:1]));     uu2 = Exact_u[:, 0:1]
    xx3 = np.hstack((X[:, -1:], T[:, -1:]));     uu3 = Exact_u[:, -1:]
    X_u_train = np.vstack([xx1, xx2, xx3])
    u_train = np.vstack([uu1, uu2, uu3])
    X_f_train = lb + (ub - lb) * _lhs(2, N_f)
    idx = np.random.choice(X_u_train.shape[0], N_u, replace=False)
    return x, t, X, T, Exact_u, X_star, u_star, X_u_train[idx, :], u_train[idx, :], X_f_train, ub, lb


def _no_plot(*a, **k):
    print("(plotting skipped: matplotlib/LaTeX are not part of the training hot path)")


plot_inf_cont_results = plot_ide_cont_results = plot_inf_disc_results = plot_ide_disc_results = _no_plot
