"""Data preparation for the continuous-time Burgers problems (restates 1d-burgers/burgersutil.py:27-36, 63-75, 99-131
and, for discrete-time inference, :38-60 of the reference).  The plot_* helpers write the run's artefact directory
(hp.json + fields.npz, figure only when matplotlib exists) through utils/plotting.py."""
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
    data = scipy.io.loadmat(path)
    t = data["t"].flatten()[:, None]
    x = data["x"].flatten()[:, None]
    Exact_u = np.real(data["usol"]).T
    if N_n is not None:
        # discrete-time inference (reference burgersutil.py:38-60): N_n points of the snapshot idx_t_0, the two boundary
        # positions, the target snapshot idx_t_1 and the q-stage implicit Runge-Kutta table of the upstream repository
        dt_ = t[idx_t_1] - t[idx_t_0]
        idx_x = np.random.choice(Exact_u.shape[1], N_n, replace=False)
        x_0 = x[idx_x, :]
        u_0 = Exact_u[idx_t_0:idx_t_0 + 1, idx_x].T
        u_0 = u_0 + noise * np.std(u_0) * np.random.randn(u_0.shape[0], u_0.shape[1])
        x_1 = np.vstack((lb, ub))
        irk_w, irk_t = load_irk(q)
        return x, t, dt_, Exact_u, x_0, u_0, x_1, x, Exact_u[idx_t_1, :], irk_w, irk_t
    X, T = np.meshgrid(x, t)
    X_star = np.hstack((X.flatten()[:, None], T.flatten()[:, None]))
    u_star = Exact_u.flatten()[:, None]
    # identification: N_u random grid points, `noise` ignored exactly like the reference (burgersutil.py:72-75)
    idx = np.random.choice(X_star.shape[0], N_u, replace=False)
    X_u_train, u_train = X_star[idx, :], u_star[idx, :]
    lb, ub = X_star.min(axis=0), X_star.max(axis=0)
    if N_0 is not None and N_1 is not None:
        # discrete-time identification (reference burgersutil.py:76-97; its np.asscalar no longer exists in numpy, .item() is
        # the same value): N_0 / N_1 points of the snapshots idx_t_0 / idx_t_1, q from the time step, IRK table split in
        # alpha (q x q) and beta (1 x q)
        Exact_xt = Exact_u.T
        idx_x = np.random.choice(Exact_xt.shape[0], N_0, replace=False)
        x_0 = x[idx_x, :]
        u_0 = Exact_xt[idx_x, idx_t_0][:, None]
        u_0 = u_0 + noise * np.std(u_0) * np.random.randn(u_0.shape[0], u_0.shape[1])
        idx_x = np.random.choice(Exact_xt.shape[0], N_1, replace=False)
        x_1 = x[idx_x, :]
        u_1 = Exact_xt[idx_x, idx_t_1][:, None]
        u_1 = u_1 + noise * np.std(u_1) * np.random.randn(u_1.shape[0], u_1.shape[1])
        dt_ = (t[idx_t_1] - t[idx_t_0]).item()
        q = int(np.ceil(0.5 * np.log(np.finfo(float).eps) / np.log(dt_)))
        weights, _ = load_irk(q)
        return x_0, u_0, x_1, u_1, x, t, dt_, q, Exact_xt, weights[0:-1, :], weights[-1:, :]
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


def load_irk(q, utils_path=None):
    """Butcher table of the q-stage Gauss-Legendre IRK scheme as shipped by the upstream PINNs repository
    (PINNs/Utilities/IRK_weights/Butcher_IRK<q>.txt), rounded to float32 like the reference does."""
    import os
    utils_path = utils_path or os.path.join(".", "PINNs", "Utilities")
    tmp = np.float32(np.loadtxt(os.path.join(utils_path, "IRK_weights", "Butcher_IRK%d.txt" % q), ndmin=2))
    return np.reshape(tmp[0:q ** 2 + q], (q + 1, q)), tmp[q ** 2 + q:]


def _grid(X_star, values, X, T):
    """Prediction on the (t, x) grid the figure shows: X_star is that grid flattened (prep_data), so a reshape; scattered
    inputs go through scipy's cubic griddata like the reference (burgersutil.py:136)."""
    values = np.asarray(values).reshape(-1)
    if values.size == X.size:
        return values.reshape(X.shape)
    from scipy.interpolate import griddata
    return griddata(X_star, values, (X, T), method="cubic")


def _finish(save_path, save_hp, draw, **fields):
    """Common tail of the plot_* helpers: stage the arrays, draw if matplotlib exists, write the artefact directory."""
    import plotting
    plotting.stage_fields(**fields)
    if plotting.have_matplotlib():
        try:
            draw(plotting)
        except Exception as exc:            # a cosmetic failure must not lose the run's numbers
            print("(figure skipped: %s)" % exc)
    if save_path is not None and save_hp is not None:
        return plotting.saveResultDir(save_path, save_hp)
    print("(no save_path/save_hp given and no interactive display: nothing written)")
    return None


def _slices(pl, x, exact_rows, pred_rows, titles):
    fig, _ = pl.newfig(1.0, 1.1)
    for k, (e, p_, ttl) in enumerate(zip(exact_rows, pred_rows, titles)):
        ax = fig.add_subplot(1, len(titles), k + 1)
        ax.plot(x, e, "b-", linewidth=2, label="Exact")
        if p_ is not None:
            ax.plot(x, p_, "r--", linewidth=2, label="Prediction")
        ax.set_xlabel("x"); ax.set_title(ttl, fontsize=10)
    return fig


def plot_inf_cont_results(X_star, u_pred, X_u_train, u_train, Exact_u, X, T, x, t, save_path=None, save_hp=None):
    U_pred = _grid(X_star, u_pred, X, T)
    rows = (25, 50, 75)                                   # the three time slices of the reference figure (burgersutil.py:166-196)
    return _finish(save_path, save_hp,
                   lambda pl: _slices(pl, x, [Exact_u[r] for r in rows], [U_pred[r] for r in rows], ["t = %.2f" % t[r, 0] for r in rows]),
                   U_pred=U_pred, Exact_u=Exact_u, x=x, t=t, X_u_train=X_u_train, u_train=u_train, slice_rows=np.array(rows),
                   rel_l2_error=np.linalg.norm(Exact_u - U_pred) / np.linalg.norm(Exact_u))


def plot_ide_cont_results(X_star, u_pred, X_u_train, u_train, Exact_u, X, T, x, t, lambda_1_value, lambda_1_value_noisy,
                          lambda_2_value, lambda_2_value_noisy, save_path=None, save_hp=None):
    U_pred = _grid(X_star, u_pred, X, T)
    rows = (25, 50, 75)
    return _finish(save_path, save_hp,
                   lambda pl: _slices(pl, x, [Exact_u[r] for r in rows], [U_pred[r] for r in rows], ["t = %.2f" % t[r, 0] for r in rows]),
                   U_pred=U_pred, Exact_u=Exact_u, x=x, t=t, X_u_train=X_u_train, u_train=u_train, slice_rows=np.array(rows),
                   lambdas=np.array([lambda_1_value, lambda_2_value], dtype=float),
                   lambdas_noisy=np.array([lambda_1_value_noisy, lambda_2_value_noisy], dtype=float),
                   lambdas_exact=np.array([1.0, 0.01 / np.pi]))


def plot_inf_disc_results(x_star, idx_t_0, idx_t_1, x_0, u_0, ub, lb, u_1_pred, Exact_u, x, t, save_path=None, save_hp=None):
    u_1_pred = np.asarray(u_1_pred).reshape(-1)
    return _finish(save_path, save_hp,
                   lambda pl: _slices(pl, x, [Exact_u[idx_t_0], Exact_u[idx_t_1]], [None, u_1_pred],
                                      ["t = %.2f" % t[idx_t_0, 0], "t = %.2f" % t[idx_t_1, 0]]),
                   u_1_pred=u_1_pred, Exact_u=Exact_u, x=x, t=t, x_star=x_star, x_0=x_0, u_0=u_0, lb=lb, ub=ub,
                   idx_t=np.array([idx_t_0, idx_t_1]),
                   rel_l2_error=np.linalg.norm(Exact_u[idx_t_1] - u_1_pred) / np.linalg.norm(Exact_u[idx_t_1]))


def plot_ide_disc_results(x_star, t_star, idx_t_0, idx_t_1, x_0, u_0, x_1, u_1, ub, lb, u_1_pred, Exact, lambda_1_value,
                          lambda_1_value_noisy, lambda_2_value, lambda_2_value_noisy, x, t, save_path=None, save_hp=None):
    """Artefacts of the discrete-time identification run (reference figure: burgersutil.py:263-330): both data snapshots, the
    exact field (x by t), the identified coefficients with and without noise."""
    return _finish(save_path, save_hp,
                   lambda pl: _slices(pl, x_star, [Exact[:, idx_t_0], Exact[:, idx_t_1]], [None, None],
                                      ["t = %.2f" % t_star[idx_t_0, 0], "t = %.2f" % t_star[idx_t_1, 0]]),
                   Exact=Exact, x_star=x_star, t_star=t_star, x_0=x_0, u_0=u_0, x_1=x_1, u_1=u_1, lb=lb, ub=ub, U_1_pred=u_1_pred,
                   idx_t=np.array([idx_t_0, idx_t_1]), lambdas=np.array([lambda_1_value, lambda_2_value], dtype=float),
                   lambdas_noisy=np.array([lambda_1_value_noisy, lambda_2_value_noisy], dtype=float),
                   lambdas_exact=np.array([1.0, 0.01 / np.pi]))
