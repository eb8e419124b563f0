"""GPU: the generic fused kernel (any hp["layers"], SURVEY quirk Q7 aside) against the numpy Taylor oracle, and as an
independent cross-check of the two specialised DMMA kernels (PINN_FORCE_GENERIC=1)."""
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def cabi():
    import pinn_cabi
    pinn_cabi.load()
    return pinn_cabi


def rand_w(layers, seed):
    from oracle import reference_port as rp
    rng = np.random.default_rng(seed)
    return rp.glorot_normal_flat(layers, rng) + 0.05 * rng.standard_normal(rp.num_params(layers))


@pytest.mark.parametrize("layers,n_f,n_u", [([2, 10, 10, 1], 333, 17), ([2, 37, 5, 64, 1], 1000, 100),
                                             ([2] + [8] * 12 + [1], 77, 5), ([2, 128, 1], 2000, 0), ([2, 20, 20, 20, 1], 1, 1),
                                             ([2] + [40] * 8 + [1], 1500, 60), ([2, 50, 50, 50, 1], 90, 9), ([2, 3, 7, 1], 50, 4)])
def test_burgers_any_layers(cabi, layers, n_f, n_u):
    from oracle import taylor as ty
    rng = np.random.default_rng(len(layers) * 100 + n_f)
    lb, ub = np.array([-1.0, 0.0]), np.array([1.0, 0.99])
    X_f = lb + (ub - lb) * rng.random((n_f, 2)); X_u = lb + (ub - lb) * rng.random((n_u, 2)); u = rng.uniform(-1, 1, (n_u, 1))
    w = rand_w(layers, 3)
    p = cabi.Pinn(cabi.BURGERS_INF, layers, lb, ub)
    assert p.kernel_info()["block"] == 256 and p.kernel_info()["dyn_smem"] == 0        # the generic kernel
    p.set_pde_params([0.01 / np.pi]); p.set_collocation(X_f[:, 0], X_f[:, 1]); p.set_data(X_u, u)
    loss, grad, parts = p.loss_grad(w=w)
    if n_u:
        f2, g2, pr = ty.burgers_loss_grad(w, layers, lb, ub, X_f, X_u, u, nu=0.01 / np.pi)
    else:
        (U, Ux, Ut, Uxx), st = ty.forward(w, layers, lb, ub, X_f)
        f = Ut + U * Ux - 0.01 / np.pi * Uxx
        c = 2 * f / n_f
        f2, g2 = float(np.sum(f * f) / n_f), ty.backward(w, layers, st, (c * Ux, c * U, c, -c * 0.01 / np.pi))
    assert abs(loss - f2) <= 1e-10 * abs(f2) and rel(grad, g2) < 1e-10
    # identification on the same net
    wi = np.concatenate([w, [0.4, -5.0]])
    q = cabi.Pinn(cabi.BURGERS_IDE, layers, lb, ub)
    if n_u:
        q.set_data(X_u, u)
        li, gi, _ = q.loss_grad(w=wi)
        f3, g3, _ = ty.burgers_loss_grad(wi, layers, lb, ub, None, X_u, u, identification=True)
        assert abs(li - f3) <= 1e-10 * abs(f3) and rel(gi, g3) < 1e-10


@pytest.mark.parametrize("layers", [[2, 30, 20, 2], [2, 100, 100, 2], [2, 16, 16, 16, 16, 16, 2]])
def test_schrodinger_any_layers(cabi, layers):
    from oracle import taylor as ty
    g = load_golden("nls_inf")
    w = rand_w(layers, 11)
    for X0 in (g["x0"], np.concatenate([g["x0"], 0 * g["x0"]], 1), g["x0"][:7]):       # quirk Q1, intended, odd count
        uv0 = g["uv0"][: X0.shape[0]]
        p = cabi.Pinn(cabi.NLS_INF, layers, g["lb"], g["ub"])
        p.set_collocation(g["X_f"][:, 0], g["X_f"][:, 1]); p.set_boundary(g["tb"]); p.set_data(X0, uv0)
        loss, grad, parts = p.loss_grad(w=w)
        f2, g2, pr = ty.schrodinger_loss_grad(w, layers, g["lb"], g["ub"], g["X_f"], g["tb"], X0, uv0)
        assert abs(loss - f2) <= 1e-10 * abs(f2) and rel(grad, g2) < 1e-10
        assert np.allclose(parts, pr, rtol=1e-10)


def test_generic_cross_checks_the_specialised_kernels(cabi):
    gb, gn = load_golden("burgers_inf"), load_golden("nls_inf")
    os.environ["PINN_FORCE_GENERIC"] = "1"
    try:
        pg = cabi.Pinn(cabi.BURGERS_INF, [2] + [20] * 8 + [1], gb["lb"], gb["ub"])
        ng = cabi.Pinn(cabi.NLS_INF, [2, 100, 100, 100, 100, 2], gn["lb"], gn["ub"])
    finally:
        del os.environ["PINN_FORCE_GENERIC"]
    ps = cabi.Pinn(cabi.BURGERS_INF, [2] + [20] * 8 + [1], gb["lb"], gb["ub"])
    ns = cabi.Pinn(cabi.NLS_INF, [2, 100, 100, 100, 100, 2], gn["lb"], gn["ub"])
    assert pg.kernel_info()["dyn_smem"] == 0 and ps.kernel_info()["dyn_smem"] > 100000
    for p in (pg, ps):
        p.set_pde_params([float(gb["nu"])]); p.set_collocation(gb["X_f"][:, 0], gb["X_f"][:, 1]); p.set_data(gb["X_u"], gb["u"])
    for p in (ng, ns):
        p.set_collocation(gn["X_f"][:, 0], gn["X_f"][:, 1]); p.set_boundary(gn["tb"]); p.set_data(gn["x0"], gn["uv0"])
    l1, g1, _ = pg.loss_grad(w=gb["w"]); l2, g2, _ = ps.loss_grad(w=gb["w"])
    assert abs(l1 - l2) <= 1e-12 * abs(l2) and rel(g1, g2) < 1e-12
    assert abs(l1 - gb["loss"]) <= 1e-10 * abs(gb["loss"])
    l1, g1, _ = ng.loss_grad(w=gn["w"]); l2, g2, _ = ns.loss_grad(w=gn["w"])
    assert abs(l1 - l2) <= 1e-12 * abs(l2) and rel(g1, g2) < 1e-12


def test_generic_net_trains_like_the_oracle(cabi):
    """Adam (TF-2.0 semantics) + device L-BFGS on a non-benchmark layer list."""
    from oracle import reference_port as rp
    g = load_golden("burgers_inf")
    layers = [2, 24, 12, 24, 1]
    w = rand_w(layers, 5)
    p = cabi.Pinn(cabi.BURGERS_INF, layers, g["lb"], g["ub"])
    p.set_pde_params([float(g["nu"])]); p.set_collocation(g["X_f"][:, 0], g["X_f"][:, 1]); p.set_data(g["X_u"], g["u"])
    p.set_weights(w)
    pb = rp.BurgersInference(layers, g["lb"], g["ub"], float(g["nu"]), g["X_f"], g["X_u"], g["u"])
    wa, la, _ = rp.adam_train(pb, w, 4, lr=0.01)
    losses = [p.adam_step(0.01) for _ in range(4)]
    assert rel(losses, la) < 1e-8 and rel(p.get_weights(), wa) < 1e-8
    tr = rp.lbfgs_fixed_step(lambda x: rp.loss_and_flat_grad(pb, x), wa, max_iter=5, learning_rate=0.8, n_correction=50,
                             tol_fun=np.finfo(float).eps)
    r = p.lbfgs(5, learning_rate=0.8, n_correction=50, tol_fun=np.finfo(float).eps, want_x_final=True)
    assert r["n_iter"] == tr.n_iter and rel(r["x_final"], tr.x_final) < 1e-7


def test_tensor_core_and_dfma_forms_of_the_hidden_layers_agree():
    """The generic kernel runs its hidden-to-hidden layers on DMMA.8x8x4 by default; PINN_GENERIC_DFMA=1 keeps the plain DFMA
    form.  Same inputs, both forms, upstream's 8 x 40 net and an odd-width one: loss and gradient agree to rounding."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    code = r'''
import os, sys, json, numpy as np
ROOT = %r
sys.path[:0] = [ROOT, os.path.join(ROOT, "pinns-tf2.0_b200", "utils")]
import pinn_cabi
out = {}
for name, layers, n_f in (("8x40", [2] + [40] * 8 + [1], 3001), ("odd", [2, 37, 5, 64, 13, 1], 777)):
    rng = np.random.default_rng(5)
    lb, ub = np.array([-1.0, 0.0]), np.array([1.0, 0.99])
    X_f = lb + (ub - lb) * rng.random((n_f, 2)); X_u = lb + (ub - lb) * rng.random((40, 2)); u = rng.uniform(-1, 1, (40, 1))
    P = sum(layers[i] * layers[i + 1] + layers[i + 1] for i in range(len(layers) - 1))
    w = 0.3 * rng.standard_normal(P)
    p = pinn_cabi.Pinn(pinn_cabi.BURGERS_INF, layers, lb, ub)
    p.set_pde_params([0.01 / np.pi]); p.set_collocation(X_f[:, 0], X_f[:, 1]); p.set_data(X_u, u)
    loss, grad, parts = p.loss_grad(w=w)
    out[name] = {"loss": loss, "grad": np.asarray(grad).tolist()}
print(json.dumps(out))
''' % ROOT
    res = {}
    for mode in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PINN_GENERIC_DFMA=mode), capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    for name in ("8x40", "odd"):
        a, b = res["0"][name], res["1"][name]
        assert abs(a["loss"] - b["loss"]) <= 1e-12 * abs(b["loss"]) and rel(a["grad"], b["grad"]) < 1e-12
        assert a["grad"] != b["grad"]          # two different summation orders actually ran
