"""GPU: discrete-time (q-stage implicit Runge-Kutta) Burgers inference -- 1d-burgers/inf_disc_burgers.py:49-127 -- on the
generic fused kernel: golden vectors (real upstream q=100 table), the full-size q=500 net against the numpy Taylor oracle
with a synthetic stage matrix, Adam, and the reference-style class through the Python surface."""
import os
import sys

import numpy as np
import pytest

from conftest import PKG, assert_matches_reference_run, load_golden

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


def make(cabi, g, layers=None):
    layers = layers or [int(v) for v in g["layers"]]
    p = cabi.Pinn(cabi.BURGERS_DISC, layers, g["lb"], g["ub"])
    p.set_pde_params([float(g["nu"]), float(g["dt"])])
    p.set_irk(g["IRK"].astype(np.float64))
    p.set_boundary(g["x_1"])
    p.set_data(g["x_0"], g["u_0"])
    return p


def test_golden_loss_grad_predict_adam(cabi):
    g = load_golden("burgers_disc")
    p = make(cabi, g)
    assert p.P == g["w"].size
    loss, grad, parts = p.loss_grad(w=g["w"])
    assert abs(loss - g["loss"]) <= 1e-10 * abs(g["loss"])
    assert rel(grad, g["grad"]) < 1e-10
    assert np.allclose(parts[:2], g["parts"], rtol=1e-10) and parts[2] == 0.0
    assert_matches_reference_run(loss, grad, "burgers_disc_loss", "burgers_disc_grad")
    assert rel(p.predict(g["x_star"])[:, -1], g["predict"]) < 1e-12
    losses = [p.adam_step(1e-3, eps=1e-8) for _ in range(3)]
    assert rel(losses, g["adam_losses"]) < 1e-8 and rel(p.get_weights(), g["adam_w"]) < 1e-8


@pytest.mark.parametrize("q,n,hidden", [(500, 250, [50, 50, 50]), (7, 33, [16, 8]), (32, 1, [50])])
def test_against_taylor_oracle_with_synthetic_stage_matrix(cabi, q, n, hidden):
    from oracle import reference_port as rp, taylor as ty
    rng = np.random.default_rng(q + n)
    layers = [1] + hidden + [q + 1]
    lb, ub = np.array([-1.0]), np.array([1.0])
    w = rp.glorot_normal_flat(layers, rng) + 0.02 * rng.standard_normal(rp.num_params(layers))
    x_0 = rng.uniform(-1, 1, (n, 1)); u_0 = -np.sin(np.pi * x_0); x_1 = np.array([[-1.0], [1.0]])
    IRK = rng.standard_normal((q + 1, q)) / q
    p = cabi.Pinn(cabi.BURGERS_DISC, layers, lb, ub)
    p.set_pde_params([0.01 / np.pi, 0.8]); p.set_irk(IRK); p.set_boundary(x_1); p.set_data(x_0, u_0)
    loss, grad, _ = p.loss_grad(w=w)
    f2, g2, _ = ty.burgers_disc_loss_grad(w, layers, lb, ub, x_0, u_0, x_1, 0.01 / np.pi, 0.8, IRK)
    assert abs(loss - f2) <= 1e-10 * abs(f2) and rel(grad, g2) < 1e-10


def test_reference_style_class_through_the_surface(capsys):
    from logger import Logger
    from neuralnetwork import NeuralNetwork
    import tensorflow as tf
    g = load_golden("burgers_disc")
    q = int(g["q"])

    class BurgersInformedNN(NeuralNetwork):              # 1d-burgers/inf_disc_burgers.py:49-127 (tape bodies elided)
        def __init__(self, hp, logger, dt, x_1, lb, ub, nu, IRK_weights, IRK_times):
            super().__init__(hp, logger, ub, lb)
            self.nu = nu
            self.dt = dt
            self.q = max(hp["q"], 1)
            self.IRK_weights = IRK_weights
            self.IRK_times = IRK_times
            self.x_1 = tf.convert_to_tensor(x_1, dtype=self.dtype)

        def U_0_model(self, x):
            raise AssertionError("tape body must have been replaced")

        def grad(self, x_0, u_0):
            raise AssertionError("tape body must have been replaced")

        def fit(self, x_0, u_0):
            self.dummy_x0_tf = tf.ones([x_0.shape[0], self.q], dtype=self.dtype)
            super().fit(x_0, u_0)

        def predict(self, x_star):
            return self.model(x_star)[:, -1]

    hp = {"N_n": 250, "q": q, "layers": [1, 50, 50, 50, q + 1], "tf_epochs": 3, "tf_lr": 0.001, "tf_b1": 0.9, "tf_eps": 1e-08,
          "nt_epochs": 0, "nt_lr": 0.8, "nt_ncorr": 50, "log_frequency": 1}
    pinn = BurgersInformedNN(hp, Logger(hp), np.array([float(g["dt"])]), g["x_1"], g["lb"], g["ub"], float(g["nu"]),
                             g["IRK"].astype(np.float64), None)
    pinn._w0 = g["w"].copy()
    pinn.logger.set_error_fn(lambda: 0.0)
    pinn.fit(g["x_0"], g["u_0"])
    assert rel(pinn.get_weights().numpy(), g["adam_w"]) < 1e-8
    assert pinn.predict(g["x_star"]).shape == (g["x_star"].shape[0],)
    assert "tf_epoch =      2" in capsys.readouterr().out
