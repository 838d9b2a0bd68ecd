"""Sample sharding for the two paths that partition across GPUs: psgd minibatches
and batch prediction.  pcd / pbcd are sequential in the coordinate order and stay on one GPU
("replicas only": independent fits per GPU).

One process per GPU (torchrun); torch.distributed (NCCL over NVLink on the B200 box, gloo in the
CPU tests) is the only collective provider.
"""
import contextlib

import numpy as np

_ACTIVE_GROUP = None


def enable_sharding(group=None):
    """Opt in: psgd fits (and sharded_predict) started after this call treat X, y as THIS rank's shard of
    the samples and run over all ranks of `group` (default: the world group).  Sharding is never implied
    by torch.distributed merely being initialised -- without this call every rank fits its own X."""
    import torch.distributed as dist
    global _ACTIVE_GROUP
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("enable_sharding needs an initialised torch.distributed process group")
    _ACTIVE_GROUP = group if group is not None else dist.group.WORLD
    return _ACTIVE_GROUP


def disable_sharding():
    global _ACTIVE_GROUP
    _ACTIVE_GROUP = None


@contextlib.contextmanager
def sharded(group=None):
    enable_sharding(group)
    try:
        yield
    finally:
        disable_sharding()


def active_group():
    """The process group psgd fits shard over, or None (single process, or sharding not enabled)."""
    import torch.distributed as dist
    if _ACTIVE_GROUP is None or not (dist.is_available() and dist.is_initialized()):
        return None
    return _ACTIVE_GROUP if dist.get_world_size(_ACTIVE_GROUP) > 1 else None


def broadcast_arrays(arrays, group, src=0):
    """Make rank `src`'s numpy arrays (in place) the arrays of every rank: sharded fits must start from
    identical parameters whatever each rank's random_state drew."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    for a in arrays:
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        dist.broadcast(t, src=dist.get_global_rank(group, src) if hasattr(dist, "get_global_rank") else src, group=group)
        a[...] = t.cpu().numpy().reshape(a.shape)


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


def sharded_predict(estimator, X_local, group=None, gather=True):
    """Batch prediction (base.py:52-100 -> kernels.poly_predict, kernels.py:140-153) with the samples
    sharded over the ranks: every rank scores ITS rows X_local with the replicated model -- no collective on
    the compute path.  gather=True: the per-rank outputs are concatenated in rank order on every rank
    (shards may differ in length); gather=False: the local scores only."""
    import torch
    import torch.distributed as dist
    group = group if group is not None else active_group()
    mine = np.ascontiguousarray(estimator._predict(X_local), dtype=np.float64)
    if group is None or not gather:
        return mine
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    n = torch.tensor([mine.shape[0]], dtype=torch.int64, device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    ns = [int(v.item()) for v in ns]
    cap = max(max(ns), 1)
    buf = torch.zeros(cap, dtype=torch.float64, device=dev)
    buf[: mine.shape[0]] = torch.from_numpy(mine).to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return np.concatenate([o[:m].cpu().numpy() for o, m in zip(outs, ns)])
