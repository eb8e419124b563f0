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
