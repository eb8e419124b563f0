"""L-BFGS iteration cost on BASELINE configs[1] (Burgers, N_f=100000): reference schedule lr .8, 50 corrections."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pinns-tf2.0_b200", "utils")):
    sys.path.insert(0, p)
import pinn_cabi, bench
n_f = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
X_f, X_u, u = bench.synthetic_problem(1234, n_f)
p = pinn_cabi.Pinn(pinn_cabi.BURGERS_INF, bench.LAYERS, bench.LB, bench.UB)
p.set_pde_params([bench.NU]); p.set_data(X_u, u); p.set_collocation(X_f[:, 0], X_f[:, 1]); p.set_weights(bench.init_weights())
for _ in range(100): p.adam_step(1e-3, sync=False)          # reference schedule: Adam first
p.sync()
out = {}
for se in (1, 10):
    w = p.get_weights()
    t0 = time.perf_counter()
    r = p.lbfgs(iters, learning_rate=0.8, n_correction=50, tol_fun=np.finfo(float).eps, sync_every=se)
    dt = time.perf_counter() - t0
    out[f"sync_every_{se}"] = {"ms_per_iter": dt / max(1, r["n_iter"]) * 1e3, "n_iter": r["n_iter"], "reason": r["reason_str"],
                               "pts_per_s": n_f * r["n_iter"] / dt}
    p.set_weights(w)
ms = p.time_kernel_ms(10) / 10
out["fused_kernel_ms"] = ms
print(json.dumps(out))
