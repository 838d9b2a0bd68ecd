"""Sample sharding for the two paths that partition across GPUs: psgd minibatches
and batch prediction.  pcd / pbcd are sequential in the coordinate order and stay on one GPU
("replicas only": independent fits per GPU).

One process per GPU (torchrun); torch.distributed (NCCL over NVLink on the B200 box, gloo in the
CPU tests) is the only collective provider.
"""
import numpy as np


def local_batches(n_local, batch_size_global, world):
    """Yield (b0, b1, b_global) for every minibatch of an epoch on one rank.

    Every rank holds n_local samples (equal shards).  Global minibatch m is the union over ranks
    of local rows [m*b_loc, (m+1)*b_loc) with b_loc = max(1, batch_size_global // world); this is
    the reference's minibatch sequence (psgd.py:150-198, last partial batch included) run on the
    dataset obtained by interleaving the shards in blocks of b_loc rows."""
    b_loc = max(1, int(batch_size_global) // int(world))
    b0 = 0
    while b0 < n_local:
        b1 = min(n_local, b0 + b_loc)
        yield b0, b1, (b1 - b0) * world
        b0 = b1


def interleave_shards(shards, batch_size_global):
    """Row order of the equivalent single-process dataset: used by the parity tests to replay a
    sharded run on the oracle.  `shards` is a list of per-rank row-index arrays of equal length."""
    world = len(shards)
    n_local = len(shards[0])
    order = []
    for b0, b1, _ in local_batches(n_local, batch_size_global, world):
        for r in range(world):
            order.append(np.asarray(shards[r][b0:b1]))
    return np.concatenate(order) if order else np.zeros(0, dtype=np.int64)


def shard_rows(n, rank, world):
    """Contiguous equal row ranges; the remainder n % world is dropped from the tail so that all
    ranks run the same number of minibatches (no rank can stall an all-reduce)."""
    per = n // world
    return rank * per, (rank + 1) * per


def global_sum(values, group=None):
    """All-reduce a small list of python numbers (n, nnz, loss sums)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(values)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(list(values), dtype=torch.float64, device=dev)
    dist.all_reduce(t, group=group)
    return t.tolist()


def sharded_predict(estimator, X, group=None):
    """Batch prediction with rows of X split over the ranks; every rank returns the full vector.
    No collective on the compute path -- only the final gather of the outputs."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return estimator._predict(X)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = X.shape[0]
    per = -(-n // world)
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.zeros(per, dtype=torch.float64, device=dev)
    if hi > lo:
        part = estimator._predict(X[lo:hi])
        mine[: hi - lo] = torch.from_numpy(np.ascontiguousarray(part)).to(dev)
    out = torch.empty(world * per, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out[:n].cpu().numpy()
