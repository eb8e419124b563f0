"""CPU, world_size 2 over gloo: the sharding algebra the multi-GPU path relies on.
Each rank evaluates its shard with the oracle (global 1/N_f, data term on rank 0), one all-reduce sums
[gradient | loss]; the result must equal the unsharded evaluation."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pinns-tf2.0_b200", "utils"))
    import sharding
    from oracle import taylor as ty
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "burgers_inf.npz"))
    layers = [int(v) for v in g["layers"]]
    n = g["X_f"].shape[0]
    lo, hi = sharding.shard_rows(n, rank, world)
    uid = sharding.exchange_nccl_uid(dist, rank, lambda: b"U" * 128)
    f, gr, _ = ty.burgers_loss_grad(g["w"], layers, g["lb"], g["ub"], g["X_f"][lo:hi], g["X_u"], g["u"], nu=float(g["nu"]),
                                    n_f_global=n, data_weight=sharding.data_weight(rank))
    buf = torch.from_numpy(np.concatenate([gr, [f]]))
    dist.all_reduce(buf)
    if rank == 0:
        np.save(out, buf.numpy())
    assert uid == b"U" * 128
    dist.destroy_process_group()


def test_two_rank_shards_sum_to_full(tmp_path):
    out = str(tmp_path / "r.npy")
    port = 29500 + os.getpid() % 500
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = np.load(out)
    g = load_golden("burgers_inf")
    assert abs(r[-1] - g["loss"]) <= 1e-13 * abs(g["loss"])
    assert np.linalg.norm(r[:-1] - g["grad"]) <= 1e-12 * np.linalg.norm(g["grad"])


def test_shard_rows_partition():
    import sharding
    for n in (1, 7, 100000, 2000000):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_rows(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        sharding.shard_rows(10, 2, 2)


def _worker_full(rank, world, port, out):
    """Schrodinger and identification shards + a replicated Adam step: after the one all-reduce every rank applies the same
    update to its replica, so the replicas stay bit-identical without a broadcast."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pinns-tf2.0_b200", "utils"))
    import sharding
    from oracle import taylor as ty, reference_port as rp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {}
    # Schrodinger: collocation rows sharded, initial-condition and boundary terms owned by rank 0
    g = np.load(os.path.join(ROOT, "tests", "golden", "nls_inf.npz"))
    layers = [int(v) for v in g["layers"]]
    n = g["X_f"].shape[0]
    lo, hi = sharding.shard_rows(n, rank, world)
    f, gr, parts = ty.schrodinger_loss_grad(g["w"], layers, g["lb"], g["ub"], g["X_f"][lo:hi], g["tb"], g["x0"], g["uv0"],
                                            n_f_global=n, aux_weight=sharding.data_weight(rank))
    buf = torch.from_numpy(np.concatenate([gr, parts]))                 # [gradient P | 3 loss parts], the library's exchange vector
    dist.all_reduce(buf)
    res["nls"] = buf.numpy().copy()
    # one replicated Adam step from the reduced gradient, on every rank
    st = rp.adam_init(gr.size)
    w1 = rp.adam_update(g["w"].copy(), buf.numpy()[:-3], st, lr=0.05, b1=0.99, eps=0.1)
    w1 = w1[0] if isinstance(w1, tuple) else w1
    gathered = [torch.zeros_like(torch.from_numpy(w1)) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(w1))
    res["replicas_equal"] = all(torch.equal(gathered[0], t) for t in gathered)
    res["w1"] = w1
    if rank == 0:
        np.savez(out, **res)
    dist.destroy_process_group()


def test_two_rank_schrodinger_shards_and_replicated_adam(tmp_path):
    out = str(tmp_path / "r.npz")
    port = 29800 + os.getpid() % 150
    mp.spawn(_worker_full, args=(2, port, out), nprocs=2, join=True)
    r = np.load(out)
    g = load_golden("nls_inf")
    assert np.linalg.norm(r["nls"][:-3] - g["grad_q1"]) <= 1e-12 * np.linalg.norm(g["grad_q1"])
    assert np.allclose(r["nls"][-3:], g["parts_q1"], rtol=1e-12) and abs(r["nls"][-3:].sum() - g["loss_q1"]) <= 1e-12 * g["loss_q1"]
    assert bool(r["replicas_equal"])
    # the replicated step equals the first step of the single-process trajectory up to the summation order of the shards
    from oracle import reference_port as rp
    st = rp.adam_init(g["w"].size)
    w_single = rp.adam_update(g["w"].copy(), g["grad_q1"], st, lr=0.05, b1=0.99, eps=0.1)
    w_single = w_single[0] if isinstance(w_single, tuple) else w_single
    assert np.linalg.norm(r["w1"] - w_single) <= 1e-12 * np.linalg.norm(w_single)
