"""TEST-ONLY stand-in for ``pinn_cabi.Pinn`` built on the CPU oracle (oracle/taylor.py, oracle/reference_port.py).

The product has no CPU path and never imports this: tests that want to exercise the HOST-side mirror (fit loops, logger
traffic, flat-parameter bridge, L-BFGS wrapper, PDE recognition) without a GPU replace the class object ``pinn_cabi.Pinn`` by
``OraclePinn`` in the test process (monkeypatch).  Method names, arguments and return shapes follow utils/pinn_cabi.py."""
import numpy as np

from oracle import reference_port as rp
from oracle import taylor as ty

BURGERS_INF, BURGERS_IDE, NLS_INF, BURGERS_DISC = 0, 1, 2, 3
_REASON = {"max iterations": 1, "max evaluations": 2, "optimality": 3, "step below tolX": 4, "f change below tolX": 5,
           "no progress along direction": 6, "initial optimality": 7}
_REASON_STR = {v: k for k, v in _REASON.items()}


class OraclePinn(object):
    created = []

    def __init__(self, pde, layers, lb, ub, device=0, rank=0, world=1, nccl_uid=None):
        self.pde, self.layers = int(pde), [int(v) for v in layers]
        self.lb, self.ub = np.asarray(lb, float).reshape(-1), np.asarray(ub, float).reshape(-1)
        self.P = rp.num_params(self.layers) + (2 if self.pde == BURGERS_IDE else 0)
        self.out_dim = self.layers[-1]
        self.w = np.zeros(self.P)
        self.params, self.irk, self.tb = None, None, None
        self.X_f, self.n_f_global, self.X, self.u, self.weight = None, None, None, None, 1.0
        self._adam, self._last_loss, self.launches = None, None, 0
        self.calls = []
        OraclePinn.created.append(self)

    # ---- problem definition
    def set_pde_params(self, params):
        self.params = [float(v) for v in np.asarray(params).reshape(-1)]

    def get_params(self):
        if self.pde == BURGERS_IDE:
            return np.array([self.w[-2], np.exp(self.w[-1])])
        return np.array(self.params or [])

    def set_irk(self, irk):
        self.irk = np.asarray(irk, float)

    def set_collocation(self, x, t, n_global=None):
        self.X_f = np.stack([np.asarray(x, float).reshape(-1), np.asarray(t, float).reshape(-1)], 1)
        self.n_f_global = int(n_global or self.X_f.shape[0])

    def set_data(self, X, u, weight=1.0):
        X, u = np.asarray(X, float), np.asarray(u, float)
        self.X = X[:, None] if X.ndim == 1 else X
        self.u = u[:, None] if u.ndim == 1 else u
        self.weight = float(weight)
        self.calls.append(("set_data", self.X.shape, self.u.shape))

    def set_boundary(self, tb):
        self.tb = np.asarray(tb, float).reshape(-1, 1)

    def set_weights(self, w):
        w = np.asarray(w, float).reshape(-1)
        assert w.size == self.P, (w.size, self.P)
        self.w = w.copy()

    def get_weights(self):
        return self.w.copy()

    # ---- hot path
    def _eval(self, w):
        self.launches += 1
        if self.pde == BURGERS_INF:
            f, g, (a, b) = ty.burgers_loss_grad(w, self.layers, self.lb, self.ub, self.X_f, self.X, self.u, nu=self.params[0],
                                                n_f_global=self.n_f_global, data_weight=self.weight)
            return f, g, np.array([a, 0.0, b])
        if self.pde == BURGERS_IDE:
            f, g, (a, b) = ty.burgers_loss_grad(w, self.layers, self.lb, self.ub, None, self.X, self.u, identification=True)
            return f, g, np.array([a, 0.0, b])
        if self.pde == NLS_INF:
            f, g, parts = ty.schrodinger_loss_grad(w, self.layers, self.lb, self.ub, self.X_f, self.tb, self.X, self.u,
                                                   n_f_global=self.n_f_global, aux_weight=self.weight)
            return f, g, np.array(parts)
        f, g, (a, b) = ty.burgers_disc_loss_grad(w, self.layers, self.lb, self.ub, self.X, self.u, self.tb, self.params[0],
                                                 self.params[1], self.irk)
        return f, g, np.array([a, b, 0.0])

    def loss_grad(self, w=None, want_grad=True):
        if w is not None:
            self.set_weights(w)
        f, g, parts = self._eval(self.w)
        return f, (g if want_grad else None), parts

    def adam_step(self, lr, b1=0.9, b2=0.999, eps=1e-7, sync=True):
        if self._adam is None:
            self._adam = rp.adam_init(self.P)
        f, g, _ = self._eval(self.w)
        self.w = rp.adam_update(self.w, g, self._adam, lr, b1, b2, eps)
        self._last_loss = f
        self.calls.append(("adam_step", lr, b1, b2, eps, sync))
        return f if sync else None

    def adam_reset(self):
        self._adam = None

    def last_loss(self):
        return self._last_loss

    def lbfgs(self, max_iter, learning_rate=1.0, n_correction=100, tol_fun=1e-5, tol_x=1e-19, sync_every=1, log_fn=None,
              want_x_final=False):
        self.calls.append(("lbfgs", max_iter, learning_rate, n_correction, tol_fun, tol_x, sync_every))

        def opfunc(x):
            f, g, _ = self._eval(x)
            self.w = np.array(x, float)                      # the closure's set_weights (neuralnetwork.py:92-95)
            return f, g
        tr = rp.lbfgs_fixed_step(opfunc, self.w, max_iter=max_iter, learning_rate=learning_rate, n_correction=n_correction,
                                 tol_fun=tol_fun, tol_x=tol_x)
        for it, f in tr.logged:
            if log_fn:
                log_fn(it, f)
        code = _REASON[tr.stop_reason]
        return {"n_iter": tr.n_iter, "n_eval": tr.n_eval, "reason": code, "reason_str": _REASON_STR[code],
                "x_final": np.array(tr.x_final) if want_x_final else None}

    # ---- off-path
    def _net_w(self):
        return self.w[:-2] if self.pde == BURGERS_IDE else self.w

    def predict(self, X):
        X = np.asarray(X, float)
        X = X[:, None] if X.ndim == 1 else X
        return ty.forward(self._net_w(), self.layers, self.lb, self.ub, X)[0][0]

    def derivatives(self, X):
        (U, Ux, Ut, Uxx), _ = ty.forward(self._net_w(), self.layers, self.lb, self.ub, np.asarray(X, float))
        return U, Ux, Ut, Uxx

    def residual(self, n=None):
        if self.pde == NLS_INF:
            (H, _, Ht, Hxx), _ = ty.forward(self.w, self.layers, self.lb, self.ub, self.X_f)
            h2 = H[:, 0] ** 2 + H[:, 1] ** 2
            return np.stack([Ht[:, 0] + 0.5 * Hxx[:, 1] + h2 * H[:, 1], Ht[:, 1] - 0.5 * Hxx[:, 0] - h2 * H[:, 0]], 1)
        pts = self.X if self.pde == BURGERS_IDE else self.X_f
        (U, Ux, Ut, Uxx), _ = ty.forward(self._net_w(), self.layers, self.lb, self.ub, pts)
        l1, kappa = (self.w[-2], np.exp(self.w[-1])) if self.pde == BURGERS_IDE else (1.0, self.params[0])
        return Ut + l1 * U * Ux - kappa * Uxx

    def sync(self):
        return None

    def close(self):
        return None

    def launch_count(self):
        return self.launches
