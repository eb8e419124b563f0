"""GPU: fused sm_100a Schrodinger kernel (through the C ABI) against the oracle golden vectors
(1dcomplex-schrodinger/inf_cont_schrodinger.py:60-129), ragged sizes against the numpy Taylor oracle, optimisers."""
import numpy as np
import pytest

from conftest import assert_matches_reference_run, load_golden

pytestmark = pytest.mark.gpu

LAYERS = [2, 100, 100, 100, 100, 2]


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def cabi():
    import pinn_cabi
    pinn_cabi.load()
    return pinn_cabi


def make(cabi, g, X0):
    p = cabi.Pinn(cabi.NLS_INF, LAYERS, g["lb"], g["ub"])
    assert p.P == 30802
    p.set_collocation(g["X_f"][:, 0], g["X_f"][:, 1])
    p.set_boundary(g["tb"])
    p.set_data(X0, g["uv0"])
    p.set_weights(g["w"])
    return p


def test_loss_grad_parts_both_ic_modes(cabi):
    g = load_golden("nls_inf")
    x0 = g["x0"]
    for tag, X0 in (("q1", x0), ("x0t0", np.concatenate([x0, 0 * x0], 1))):
        p = make(cabi, g, X0)
        loss, grad, parts = p.loss_grad()
        assert abs(loss - g["loss_" + tag]) <= 1e-10 * abs(g["loss_" + tag])
        assert rel(grad, g["grad_" + tag]) < 1e-10
        assert np.allclose(parts, g["parts_" + tag], rtol=1e-10)      # (mse_0, mse_b, mse_f)
        assert_matches_reference_run(loss, grad, "nls_loss_" + tag, "nls_grad_" + tag)
        loss2, grad2, _ = p.loss_grad(w=g["w"])
        assert loss2 == loss and np.array_equal(grad, grad2)          # deterministic


def test_probes_predict_residual(cabi):
    g = load_golden("nls_inf")
    p = make(cabi, g, g["x0"])
    U, Ux, Ut, Uxx = p.derivatives(g["X_f"][:32])
    got = np.stack([U[:, 0], U[:, 1], Ux[:, 0], Ux[:, 1], Ut[:, 0], Ut[:, 1], Uxx[:, 0], Uxx[:, 1]], 1)
    assert rel(got, g["probes"]) < 1e-11
    assert rel(p.predict(g["X_star"]), g["predict"]) < 1e-12
    assert rel(p.residual(g["X_f"].shape[0]), g["residual"]) < 1e-10


def test_adam_trajectory_reference_hyperparameters(cabi):
    g = load_golden("nls_inf")
    p = make(cabi, g, g["x0"])
    losses = [p.adam_step(0.05, b1=0.99, eps=0.1) for _ in range(3)]      # inf_cont_schrodinger.py:33-36
    assert rel(losses, g["adam_losses"]) < 1e-8
    assert rel(p.get_weights(), g["adam_w"]) < 1e-8


@pytest.mark.parametrize("n_f,n_0,n_b", [(1, 1, 1), (15, 3, 2), (17, 50, 50), (2400, 7, 0), (5000, 0, 50), (20000, 50, 50)])
def test_ragged_sizes_against_taylor_oracle(cabi, n_f, n_0, n_b):
    from oracle import taylor as ty
    g = load_golden("nls_inf")
    rng = np.random.default_rng(n_f + 31 * n_0 + 7 * n_b)
    lb, ub = g["lb"], g["ub"]
    X_f = lb + (ub - lb) * rng.random((n_f, 2))
    tb = rng.uniform(0, ub[1], (n_b, 1))
    x0 = rng.uniform(-5, 5, (n_0, 1)); uv0 = rng.standard_normal((n_0, 2))
    p = cabi.Pinn(cabi.NLS_INF, LAYERS, lb, ub)
    p.set_collocation(X_f[:, 0], X_f[:, 1]); p.set_boundary(tb); p.set_data(x0, uv0)
    loss, grad, parts = p.loss_grad(w=g["w"])
    # the oracle divides by n_0 / n_b; empty sets contribute nothing on either side
    f2, g2, pr = ty.schrodinger_loss_grad(g["w"], LAYERS, lb, ub, X_f, tb if n_b else np.zeros((0, 1)),
                                          x0 if n_0 else np.zeros((0, 1)), uv0 if n_0 else np.zeros((0, 2))) \
        if (n_0 and n_b) else _oracle_partial(ty, g["w"], lb, ub, X_f, tb, x0, uv0)
    assert abs(loss - f2) <= 1e-10 * abs(f2)
    assert rel(grad, g2) < 1e-10


def _oracle_partial(ty, w, lb, ub, X_f, tb, x0, uv0):
    """Oracle with an empty initial-condition or boundary set (numpy means of empty arrays are NaN, so assemble by hand)."""
    big = ty.schrodinger_loss_grad
    n_f = X_f.shape[0]
    one_t = np.array([[0.3]]); one_x = np.array([[0.1]]); one_u = np.zeros((1, 2))
    f_all, g_all, parts = big(w, LAYERS, lb, ub, X_f, tb if tb.shape[0] else one_t, x0 if x0.shape[0] else one_x,
                              uv0 if x0.shape[0] else one_u)
    # subtract the dummy term that was added for the empty set
    f_col, g_col, pc = big(w, LAYERS, lb, ub, X_f[:0], tb if tb.shape[0] else one_t, x0 if x0.shape[0] else one_x,
                           uv0 if x0.shape[0] else one_u, n_f_global=n_f)
    f = parts[2]; g = g_all - g_col
    if tb.shape[0]:
        fb, gb, pb = big(w, LAYERS, lb, ub, X_f[:0], tb, one_x, one_u, n_f_global=n_f, aux_weight=1.0)
        f0, g0, p0 = big(w, LAYERS, lb, ub, X_f[:0], tb, one_x, one_u, n_f_global=n_f, aux_weight=0.0)
        # boundary term alone = (aux=1) - (ic dummy): isolate via the parts
        f += pb[1]
        gic_dummy = _ic_only(ty, w, lb, ub, one_x, one_u)
        g += gb - gic_dummy
    if x0.shape[0]:
        gic = _ic_only(ty, w, lb, ub, x0, uv0)
        (H0, _, _, _), _ = ty.forward(w, LAYERS, lb, ub, x0)
        f += float(np.sum((H0 - uv0) ** 2) / x0.shape[0])
        g += gic
    return f, g, None


def _ic_only(ty, w, lb, ub, x0, uv0):
    (H0, _, _, _), st0 = ty.forward(w, LAYERS, lb, ub, x0)
    r0 = H0 - uv0
    z = np.zeros_like(r0)
    return ty.backward(w, LAYERS, st0, (2.0 * r0 / r0.shape[0], z, z, z))


def test_lbfgs_short_run(cabi):
    """Three device L-BFGS iterations on the NLS problem against the oracle's control flow driven by the Taylor oracle."""
    from oracle import reference_port as rp, taylor as ty
    g = load_golden("nls_inf")
    X_f = g["X_f"][:96]
    p = cabi.Pinn(cabi.NLS_INF, LAYERS, g["lb"], g["ub"])
    p.set_collocation(X_f[:, 0], X_f[:, 1]); p.set_boundary(g["tb"]); p.set_data(g["x0"], g["uv0"]); p.set_weights(g["w"])
    op = lambda x: ty.schrodinger_loss_grad(x, LAYERS, g["lb"], g["ub"], X_f, g["tb"], g["x0"], g["uv0"])[:2]
    tr = rp.lbfgs_fixed_step(op, g["w"], max_iter=3, learning_rate=1.2, n_correction=50, tol_fun=np.finfo(float).eps)
    r = p.lbfgs(3, learning_rate=1.2, n_correction=50, tol_fun=np.finfo(float).eps, want_x_final=True)
    assert r["n_iter"] == tr.n_iter and r["n_eval"] == tr.n_eval
    assert rel(r["x_final"], tr.x_final) < 1e-8
    assert rel(p.get_weights(), tr.x_eval[-1]) < 1e-8


def test_lbfgs_twenty_iterations_at_full_size(cabi):
    """BASELINE configs[2] size (N_f = 20 000, the lbfgs_iterate<32,1024> instantiation with P = 30 802): twenty device
    iterations against the oracle's control flow driven by the numpy Taylor oracle.  Rounding differences between two fp64
    implementations of a fixed-step L-BFGS grow about tenfold per ten iterations, hence 1e-6 on the trajectory."""
    from oracle import reference_port as rp, taylor as ty
    g = load_golden("nls_inf")
    rng = np.random.default_rng(20)
    X_f = g["lb"] + (g["ub"] - g["lb"]) * rng.random((20000, 2))
    p = cabi.Pinn(cabi.NLS_INF, LAYERS, g["lb"], g["ub"])
    p.set_collocation(X_f[:, 0], X_f[:, 1]); p.set_boundary(g["tb"]); p.set_data(g["x0"], g["uv0"]); p.set_weights(g["w"])
    op = lambda x: ty.schrodinger_loss_grad(x, LAYERS, g["lb"], g["ub"], X_f, g["tb"], g["x0"], g["uv0"])[:2]
    tr = rp.lbfgs_fixed_step(op, g["w"], max_iter=20, learning_rate=0.8, n_correction=50, tol_fun=np.finfo(float).eps)
    logged = []
    r = p.lbfgs(20, learning_rate=0.8, n_correction=50, tol_fun=np.finfo(float).eps, sync_every=7, want_x_final=True,
                log_fn=lambda it, f: logged.append((it, f)))
    assert r["n_iter"] == tr.n_iter == 20 and r["n_eval"] == tr.n_eval
    assert rel(r["f_hist"], tr.f_hist) < 1e-6
    assert [it for it, _ in logged] == [it for it, _ in tr.logged]
    assert rel(r["x_final"], tr.x_final) < 1e-6 and rel(p.get_weights(), tr.x_eval[-1]) < 1e-6
