"""Times the generic fused kernel (hidden layers on DMMA; PINN_GENERIC_DFMA=1: plain DFMA) on a few layer lists.  Run from the repo root."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "pinns-tf2.0_b200", "utils"))
import pinn_cabi
from bench import synthetic_problem, init_weights, LB, UB, NU, ADAM_LR, timed_adam_steps
out = {}
for name, L, n in [("8x40", [2] + [40] * 8 + [1], 20000), ("8x40_100k", [2] + [40] * 8 + [1], 100000), ("4x100", [2] + [100] * 4 + [1], 20000),
                   ("8x20g", [2] + [20] * 8 + [1], 100000), ("3x50", [2, 50, 50, 50, 1], 20000)]:
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    if name == "8x20g":
        os.environ["PINN_FORCE_GENERIC"] = "1"
    X_f, X_u, u = synthetic_problem(40, n)
    p = pinn_cabi.Pinn(pinn_cabi.BURGERS_INF, L, LB, UB)
    os.environ.pop("PINN_FORCE_GENERIC", None)
    p.set_pde_params([NU]); p.set_data(X_u, u); p.set_collocation(X_f[:, 0], X_f[:, 1]); p.set_weights(init_weights(L))
    for _ in range(3):
        p.adam_step(ADAM_LR, sync=False)
    p.sync()
    ms = float(np.mean(timed_adam_steps(p, 10, flush=False)))
    k_ms = p.time_kernel_ms(5) / 5
    S = sum(L[i] * L[i + 1] for i in range(len(L) - 1))
    out[name] = {"ms_step": ms, "kernel_ms": k_ms, "frac": n * 24.0 * S / (k_ms * 1e-3) / 1e12 / 37.0, "info": p.kernel_info()}
    p.close()
    print(name, out[name], flush=True)
