"""Small driver for ncu captures: a few launches of the fused Burgers kernel at BASELINE configs[1] size."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pinns-tf2.0_b200", "utils")):
    sys.path.insert(0, p)
import pinn_cabi
import bench
n_f = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
X_f, X_u, u = bench.synthetic_problem(1234, n_f)
p = pinn_cabi.Pinn(pinn_cabi.BURGERS_INF, bench.LAYERS, bench.LB, bench.UB)
p.set_pde_params([bench.NU]); p.set_data(X_u, u); p.set_collocation(X_f[:, 0], X_f[:, 1]); p.set_weights(bench.init_weights())
for _ in range(iters):
    loss, _, _ = p.loss_grad()
print("loss", loss)
