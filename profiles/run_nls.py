"""Timing / ncu driver for the fused NLS kernel at BASELINE configs[2] size (N_f=20000, N_0=N_b=50)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pinns-tf2.0_b200", "utils")):
    sys.path.insert(0, p)
import pinn_cabi
from neuralnetwork import _glorot_normal
LAYERS = [2, 100, 100, 100, 100, 2]
n_f = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cpu = "--cpu" in sys.argv
rng = np.random.default_rng(1234)
lb, ub = np.array([-5.0, 0.0]), np.array([5.0, np.pi / 2])
X_f = lb + (ub - lb) * rng.random((n_f, 2)); tb = rng.uniform(0, ub[1], (50, 1)); x0 = rng.uniform(-5, 5, (50, 1))
uv0 = np.stack([2 / np.cosh(x0[:, 0]), 0 * x0[:, 0]], 1)
w = _glorot_normal(LAYERS, np.random.default_rng(1234))
p = pinn_cabi.Pinn(pinn_cabi.NLS_INF, LAYERS, lb, ub)
p.set_collocation(X_f[:, 0], X_f[:, 1]); p.set_boundary(tb); p.set_data(x0, uv0); p.set_weights(w)
loss, _, parts = p.loss_grad()
p.time_kernel_ms(2)
ms = p.time_kernel_ms(iters) / iters
S = 2 * 100 + 3 * 10000 + 200
flops = n_f * 24 * S + 150 * 24 * S
# Adam step timing
for _ in range(3): p.adam_step(0.05, 0.99, 0.999, 0.1, sync=False)
p.sync(); t0 = time.perf_counter()
for _ in range(iters): p.adam_step(0.05, 0.99, 0.999, 0.1, sync=False)
p.sync(); step_ms = (time.perf_counter() - t0) / iters * 1e3
out = {"n_f": n_f, "kernel_ms": ms, "adam_step_ms": step_ms, "pts_per_s_step": n_f / step_ms * 1e3, "tflops_alg": flops / ms / 1e9,
       "frac_of_37TF": flops / ms / 1e9 / 36.99, "loss": loss, "kernel_info": p.kernel_info()}
if cpu:
    from oracle import reference_port as rp
    pb = rp.SchrodingerInference(LAYERS, lb, ub, X_f, tb, x0, uv0)
    rp.loss_and_flat_grad(pb, w); t0 = time.perf_counter()
    for _ in range(3): rp.loss_and_flat_grad(pb, w)
    sec = (time.perf_counter() - t0) / 3
    import torch
    out["cpu_port_pts_per_s"] = n_f / sec; out["cpu_cores"] = torch.get_num_threads()
print(json.dumps(out))
