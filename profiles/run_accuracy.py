"""End-of-training accuracy (BASELINE metric, second half): 1d-burgers/inf_cont_burgers.py default schedule
(N_u=100, N_f=10000, 100 Adam @0.03 then 200 L-BFGS @0.8, 50 corrections) through the reference-style Python surface
on the B200 core, same data and initial weights as the CPU oracle run stored in tests/golden/burgers_accuracy.npz."""
import io, json, os, sys, time, contextlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pinns-tf2.0_b200", "utils")):
    sys.path.insert(0, p)
from logger import Logger
from neuralnetwork import NeuralNetwork
g = np.load(os.path.join(ROOT, "tests", "golden", "burgers_accuracy.npz"))
X_star, u_star = g["X_star"].astype(np.float64), g["u_star"].astype(np.float64)


class BurgersInformedNN(NeuralNetwork):                       # 1d-burgers/inf_cont_burgers.py:48-56
    def __init__(self, hp, logger, X_f, ub, lb, nu):
        super().__init__(hp, logger, ub, lb)
        self.nu = nu
        self.x_f = self.tensor(X_f[:, 0:1]); self.t_f = self.tensor(X_f[:, 1:2])


def run(tf_epochs, nt_epochs):
    hp = {"N_u": 100, "N_f": 10000, "layers": [2] + [20] * 8 + [1], "tf_epochs": tf_epochs, "tf_lr": 0.03, "tf_b1": 0.9,
          "tf_eps": None, "nt_epochs": nt_epochs, "nt_lr": 0.8, "nt_ncorr": 50, "log_frequency": 50}
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        logger = Logger(hp)
        pinn = BurgersInformedNN(hp, logger, g["X_f"], g["ub"], g["lb"], nu=0.01 / np.pi)
        pinn._w0 = g["w0"].copy()
        err = lambda: float(np.linalg.norm(u_star - pinn.predict(X_star)) / np.linalg.norm(u_star))
        logger.set_error_fn(err)
        t0 = time.perf_counter()
        pinn.fit(g["X_u"], g["u"])
        dt = time.perf_counter() - t0
    loss, _ = pinn.grad(g["X_u"], g["u"])
    return {"tf_epochs": tf_epochs, "nt_epochs": nt_epochs, "train_seconds": dt, "final_loss": float(loss), "rel_l2_error_u": err(),
            "log_tail": buf.getvalue().strip().split("\n")[-3:]}


# one-time process costs (CUDA context, module load, first cudaMalloc) are paid here, outside the training times below
_t0 = time.perf_counter()
import pinn_cabi
_warm = pinn_cabi.Pinn(pinn_cabi.BURGERS_INF, [2] + [20] * 8 + [1], g["lb"], g["ub"])
_warm.set_pde_params([0.01 / np.pi]); _warm.set_collocation(g["X_f"][:64, 0], g["X_f"][:64, 1]); _warm.set_data(g["X_u"], g["u"])
_warm.adam_step(1e-3); _warm.lbfgs(2, learning_rate=0.8, n_correction=50, tol_fun=1e-300); _warm.close()
process_start_seconds = time.perf_counter() - _t0

out = {"process_start_seconds (CUDA context + module load, once per process)": process_start_seconds,
       "oracle_cpu_default_schedule": {"final_loss": float(g["oracle_lbfgs_f"][-1]), "rel_l2_error_u": float(g["oracle_error"])},
       "b200_default_schedule": run(100, 200), "b200_long_schedule": run(2000, 5000)}
print(json.dumps(out, indent=1))
