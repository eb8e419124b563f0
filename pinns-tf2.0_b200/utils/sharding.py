"""Data-parallel sharding of the collocation set (one process per GPU).

The loss is a mean over independent collocation points, so rank r evaluates rows [r*N/world, (r+1)*N/world) with the
GLOBAL 1/N_f weight, the small replicated terms (data / initial / boundary points) are owned by rank 0
(weight 1 there, 0 elsewhere), and ONE ncclAllReduce over [gradient | loss parts] makes every rank hold the full
loss and gradient; optimiser state is replicated, so no broadcast is needed.  The reference has no such path.
"""
import os


def shard_rows(n, rank, world):
    """Contiguous row block of rank `rank`: [lo, hi)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (rank * n) // world, ((rank + 1) * n) // world


def data_weight(rank):
    return 1.0 if rank == 0 else 0.0


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def exchange_nccl_uid(dist, rank, make_uid):
    """Rank 0 creates the ncclUniqueId (pinn_nccl_unique_id), everybody receives it through torch.distributed."""
    box = [make_uid() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def connect_p2p(dist, pinn, world):
    """All-gather the ranks' IPC handles through the control plane and map the peers' exchange buffers: the evaluation's
    tail then runs as ONE kernel (reduction + NVLink all-to-all push + rank-ordered sum + Adam) instead of
    reduce -> ncclAllReduce -> Adam.  Default; PINN_COLLECTIVE=nccl keeps the NCCL path.  The decision is collective: if
    the export or the mapping fails on ANY rank, every rank falls back to NCCL."""
    if world < 2 or os.environ.get("PINN_COLLECTIVE", "p2p") == "nccl":
        return False
    try:
        mine = pinn.p2p_export()
    except Exception:
        mine = None
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    ok = all(h is not None for h in handles)
    if ok:
        try:
            pinn.p2p_connect(handles)
        except Exception:
            ok = False
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    ok = all(flags)
    if not ok:
        try:
            pinn.p2p_enable(False)
        except Exception:
            pass
    dist.barrier()
    return ok
