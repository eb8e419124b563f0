"""Multi-GPU parity check (run under torchrun on N GPUs; not a pytest file):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Sharded loss/gradient/Adam/L-BFGS over NCCL must match the single-GPU evaluation of the whole set."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pinns-tf2.0_b200", "utils")):
    sys.path.insert(0, p)
import pinn_cabi  # noqa: E402
import sharding  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


def main():
    rank, local_rank, world = sharding.env_rank_world()
    torch.cuda.set_device(local_rank)
    dist.init_process_group("gloo")      # control plane only; the data-path NCCL communicator lives inside the library
    uid = sharding.exchange_nccl_uid(dist, rank, pinn_cabi.nccl_unique_id)
    layers = [2] + [20] * 8 + [1]
    rng = np.random.default_rng(5)
    lb, ub = np.array([-1.0, 0.0]), np.array([1.0, 0.99])
    n_f = 40000
    X_f = lb + (ub - lb) * rng.random((n_f, 2))
    X_u = lb + (ub - lb) * rng.random((100, 2)); u = rng.uniform(-1, 1, (100, 1))
    w = np.load(os.path.join(ROOT, "tests", "golden", "burgers_inf.npz"))["w"]
    lo, hi = sharding.shard_rows(n_f, rank, world)
    p = pinn_cabi.Pinn(pinn_cabi.BURGERS_INF, layers, lb, ub, device=local_rank, rank=rank, world=world, nccl_uid=uid)
    used_p2p = sharding.connect_p2p(dist, p, world)
    p.set_pde_params([0.01 / np.pi])
    p.set_collocation(X_f[lo:hi, 0], X_f[lo:hi, 1], n_global=n_f)
    p.set_data(X_u, u, weight=sharding.data_weight(rank))
    # single-GPU evaluation of the whole set on every rank (world = 1 handle)
    s = pinn_cabi.Pinn(pinn_cabi.BURGERS_INF, layers, lb, ub, device=local_rank)
    s.set_pde_params([0.01 / np.pi]); s.set_collocation(X_f[:, 0], X_f[:, 1]); s.set_data(X_u, u)
    l1, g1, _ = s.loss_grad(w=w)
    lN, gN, _ = p.loss_grad(w=w)
    assert abs(lN - l1) <= 1e-12 * abs(l1), (lN, l1)
    assert rel(gN, g1) < 1e-12
    s.set_weights(w); p.set_weights(w)
    for _ in range(5):
        a1 = s.adam_step(1e-3); aN = p.adam_step(1e-3)
    assert abs(aN - a1) <= 1e-10 * abs(a1) and rel(p.get_weights(), s.get_weights()) < 1e-10
    s.set_weights(w); p.set_weights(w)
    r1 = s.lbfgs(8, learning_rate=0.8, n_correction=50, tol_fun=np.finfo(float).eps, sync_every=3, want_x_final=True)
    rN = p.lbfgs(8, learning_rate=0.8, n_correction=50, tol_fun=np.finfo(float).eps, sync_every=3, want_x_final=True)
    assert rN["n_iter"] == r1["n_iter"] and rel(rN["x_final"], r1["x_final"]) < 1e-8
    # every rank holds identical weights (replicated optimiser state, no broadcast)
    wt = torch.from_numpy(p.get_weights())
    ws = [torch.empty_like(wt) for _ in range(world)]
    dist.all_gather(ws, wt)
    assert all(torch.equal(ws[0], x) for x in ws)
    if rank == 0:
        print(f"mgpu_check ok: p2p={used_p2p} world={world} loss={lN:.12e} rel_grad={rel(gN, g1):.1e} lbfgs_iters={rN['n_iter']}")
    dist.barrier()
    p.close(); s.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
