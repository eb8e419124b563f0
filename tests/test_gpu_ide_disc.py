"""GPU: discrete-time (implicit Runge-Kutta) Burgers IDENTIFICATION -- 1d-burgers/ide_disc_burgers.py:48-203 -- on the generic
fused kernel (pde id 4): golden vectors built from the script's own data (N_0 = 199, N_1 = 201, q = 81 upstream table, float32
stage matrices), the numbers the reference's own class produced, the numpy Taylor oracle at other sizes, Adam / L-BFGS
trajectories, and the reference-style class (own four-argument fit) through the Python surface."""
import os
import sys

import numpy as np
import pytest

from conftest import PKG, assert_matches_reference_run, load_golden, load_reference_run

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(PKG, "shims"))


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def cabi():
    import pinn_cabi
    pinn_cabi.load()
    return pinn_cabi


def make(cabi, g):
    p = cabi.Pinn(cabi.BURGERS_IDE_DISC, [int(v) for v in g["layers"]], g["lb"], g["ub"])
    p.set_pde_params([float(g["dt"])])
    p.set_irk(cabi.irk_ide_disc(g["IRK_alpha"], g["IRK_beta"]))          # float32 difference, like the reference (:107)
    p.set_snapshot(0, g["x_0"], g["u_0"])
    p.set_snapshot(1, g["x_1"], g["u_1"])
    return p


def test_golden_loss_grad_at_two_points(cabi):
    g = load_golden("burgers_ide_disc")
    p = make(cabi, g)
    assert p.P == g["w"].size == 9333
    for tag in ("", "2"):
        loss, grad, parts = p.loss_grad(w=g["w" + tag])
        assert abs(loss - g["loss" + tag]) <= 1e-10 * abs(g["loss" + tag])
        assert rel(grad, g["grad" + tag]) < 1e-10
        assert rel(grad[-2:], g["grad" + tag][-2:]) < 1e-10               # d/d lambda_1, d/d lambda_2
        assert np.allclose(parts[:2], g["parts" + tag], rtol=1e-10) and parts[2] == 0.0
        assert_matches_reference_run(loss, grad, "burgers_ide_disc_loss" + tag, "burgers_ide_disc_grad" + tag)
    l1, kappa = p.get_params()
    assert l1 == g["w2"][-2] and abs(kappa - np.exp(g["w2"][-1])) < 1e-15


def test_float64_difference_of_the_tables_is_a_different_model(cabi):
    """The reference forms IRK_beta - IRK_alpha in float32; a float64 difference moves the loss at the 1e-10 level, which the
    1e-10 parity bar can see -- the binding must not 'improve' on the reference here."""
    g = load_golden("burgers_ide_disc")
    p = make(cabi, g)
    a64, b64 = g["IRK_alpha"].astype(np.float64), g["IRK_beta"].astype(np.float64)
    p.set_irk(np.concatenate([a64, -(b64 - a64)], 0))
    loss, _, _ = p.loss_grad(w=g["w2"])
    assert abs(loss - g["loss2"]) > 1e-13 * abs(g["loss2"])


def test_adam_and_lbfgs_trajectories(cabi):
    g = load_golden("burgers_ide_disc")
    p = make(cabi, g)
    p.set_weights(g["w2"])
    losses = [p.adam_step(1e-3) for _ in range(3)]
    assert rel(losses, g["adam_losses"]) < 1e-8 and rel(p.get_weights(), g["adam_w"]) < 1e-8
    p = make(cabi, g)
    p.set_weights(g["w2"])
    r = p.lbfgs(4, learning_rate=0.8, n_correction=50, tol_fun=np.finfo(float).eps, sync_every=1, want_x_final=True)
    assert r["n_iter"] == 4 and r["n_eval"] == len(g["lbfgs_f"])
    assert rel(r["f_hist"], g["lbfgs_f"]) < 1e-7
    assert rel(r["x_final"], g["lbfgs_x_final"]) < 1e-7
    assert rel(p.get_weights(), g["lbfgs_x_eval"][-1]) < 1e-7             # the model keeps the last EVALUATED point


@pytest.mark.parametrize("q,n0,n1,hidden", [(81, 199, 201, [50, 50, 50]), (5, 17, 3, [12, 7]), (33, 1, 40, [20])])
def test_against_taylor_oracle(cabi, q, n0, n1, hidden):
    from oracle import reference_port as rp, taylor as ty
    rng = np.random.default_rng(q + n0)
    layers = [1] + hidden + [q]
    lb, ub = np.array([-1.0]), np.array([1.0])
    w = np.concatenate([rp.glorot_normal_flat(layers, rng) + 0.02 * rng.standard_normal(rp.num_params(layers)), [0.4, -3.0]])
    x_0 = rng.uniform(-1, 1, (n0, 1)); u_0 = -np.sin(np.pi * x_0)
    x_1 = rng.uniform(-1, 1, (n1, 1)); u_1 = -0.5 * np.sin(np.pi * x_1)
    alpha = (rng.standard_normal((q, q)) / q).astype(np.float32); beta = (rng.standard_normal((1, q)) / q).astype(np.float32)
    p = cabi.Pinn(cabi.BURGERS_IDE_DISC, layers, lb, ub)
    p.set_pde_params([0.8]); p.set_irk(cabi.irk_ide_disc(alpha, beta)); p.set_snapshot(0, x_0, u_0); p.set_snapshot(1, x_1, u_1)
    loss, grad, parts = p.loss_grad(w=w)
    f2, g2, parts2 = ty.burgers_ide_disc_loss_grad(w, layers, lb, ub, x_0, u_0, x_1, u_1, 0.8, alpha, beta)
    assert abs(loss - f2) <= 1e-10 * abs(f2) and rel(grad, g2) < 1e-10
    assert np.allclose(parts[:2], parts2, rtol=1e-10)


def test_reference_style_class_through_the_surface(capsys):
    """The class statement of ide_disc_burgers.py (tape bodies elided; its own four-argument fit/predict kept by NAME): fit() runs
    the native loop with the reference's initial lambdas, predict() returns (U_0, U_1), get_params() the identified pair."""
    from logger import Logger
    from neuralnetwork import NeuralNetwork
    g, r = load_golden("burgers_ide_disc"), load_reference_run()
    q = int(g["q"])

    class BurgersInformedNN(NeuralNetwork):              # 1d-burgers/ide_disc_burgers.py:48-203
        def __init__(self, hp, logger, dt, lb, ub, q, IRK_alpha, IRK_beta):
            super().__init__(hp, logger, ub, lb)
            self.dt = dt
            self.q = max(q, 1)
            self.IRK_alpha = IRK_alpha
            self.IRK_beta = IRK_beta

        def autograd(self, U, x, dummy):
            raise AssertionError("tape body must have been replaced")

        def U_0_model(self, x, customDummy=None):
            raise AssertionError("tape body must have been replaced")

        def U_1_model(self, x, customDummy=None):
            raise AssertionError("tape body must have been replaced")

        def loss(self, x_0, u_0, x_1, u_1):
            raise AssertionError("tape body must have been replaced")

        def grad(self, x_0, u_0, x_1, u_1):
            raise AssertionError("tape body must have been replaced")

        def fit(self, x_0, u_0, x_1, u_1):
            raise AssertionError("the four-argument fit must run the native loop")

        def predict(self, x_star):
            raise AssertionError("the two-snapshot predict must run natively")

    hp = {"N_0": 199, "N_1": 201, "layers": [1, 50, 50, 50, q], "tf_epochs": 3, "tf_lr": 0.001, "tf_b1": 0.9, "tf_eps": None,
          "nt_epochs": 4, "nt_lr": 0.8, "nt_ncorr": 50, "log_frequency": 1}
    pinn = BurgersInformedNN(hp, Logger(hp), float(g["dt"]), g["lb"], g["ub"], q, g["IRK_alpha"], g["IRK_beta"])
    pinn._w0 = np.asarray(r["burgers_ide_disc_fit_w0"])[:-2].copy()      # the weights the reference's own fit() started from
    pinn.logger.set_error_fn(lambda: 0.0)
    # loss / grad at the fit's starting point equal what the reference's class computed there
    loss, grads = pinn.grad(g["x_0"], g["u_0"], g["x_1"], g["u_1"])
    assert len(grads) == 2 * 4 + 2 and grads[-1].shape == (1,)
    pinn.fit(g["x_0"], g["u_0"], g["x_1"], g["u_1"])
    out = capsys.readouterr().out
    assert "tf_epoch =      2" in out and "nt_epoch =      3" in out and "l1 = " in out
    assert rel(pinn.get_weights().numpy(), r["burgers_ide_disc_fit_w"]) < 1e-7        # weights the reference's fit() left in the model
    U0, U1 = pinn.predict(g["x_star"])
    assert U0.shape == U1.shape == (g["x_star"].shape[0], q)
    l1, l2 = pinn.get_params(numpy=True)
    w = pinn.get_weights().numpy()
    assert l1 == w[-2] and abs(l2 - np.exp(w[-1])) < 1e-15


def test_predict_matches_golden(cabi):
    from neuralnetwork import _ide_disc_models
    g = load_golden("burgers_ide_disc")

    class Shell(object):
        pass
    sh = Shell()
    p = make(cabi, g)
    p.set_weights(g["w2"])
    sh._native = lambda: p
    sh.IRK_alpha, sh.IRK_beta, sh.dt = g["IRK_alpha"], g["IRK_beta"], g["dt"]
    U0, U1 = _ide_disc_models(sh, g["x_star"])
    assert rel(U0, g["predict_U0"]) < 1e-11 and rel(U1, g["predict_U1"]) < 1e-11
