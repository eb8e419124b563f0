"""Timing / ncu driver for the discrete-time Burgers model (1d-burgers/inf_disc_burgers.py: [1,50,50,50,501], q = 500, N = 250 + 2) on
the generic fused kernel.  Run from the repo root."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "pinns-tf2.0_b200", "utils"))
import pinn_cabi
from bench import init_weights, NU, timed_adam_steps
q = 500
L = [1, 50, 50, 50, q + 1]
rng = np.random.default_rng(11)
x0 = rng.uniform(-1, 1, (250, 1)); u0 = -np.sin(np.pi * x0)
p = pinn_cabi.Pinn(pinn_cabi.BURGERS_DISC, L, [-1.0], [1.0])
p.set_pde_params([NU, 0.8]); p.set_irk(rng.standard_normal((q + 1, q)) / q); p.set_boundary(np.array([-1.0, 1.0]))
p.set_data(x0, u0); p.set_weights(init_weights(L))
for _ in range(5):
    p.adam_step(1e-3, eps=1e-8, sync=False)
p.sync()
ms = float(np.mean(timed_adam_steps(p, 50, flush=False, lr=1e-3, eps=1e-8)))
print(json.dumps({"ms_per_step": ms, "kernel_ms": p.time_kernel_ms(20) / 20, "info": p.kernel_info()}))
