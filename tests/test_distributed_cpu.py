"""world_size-2 gloo test of the psgd sample-sharding logic on CPU.

Each rank owns an equal shard of the rows, computes the minibatch gradient of ITS rows (numpy
stand-in for the CUDA gradient kernel), the dense gradients are summed with all_reduce, and every
rank applies the identical update + prox.  The result must equal the single-process oracle run on
the dataset interleaved by distributed.interleave_shards with batch_size = world * b_loc."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _grad_deg2(X, y, rows, P, w, lams, loss_id, O):
    """dense (grad_P [d,k], grad_w [d], loss sum) of FM degree 2 over `rows` (psgd.py:47-91)."""
    d, k = P.shape
    gP, gw, ls = np.zeros((d, k)), np.zeros(d), 0.0
    L = O.lib()
    for i in rows:
        x = X[i]
        a1 = x @ P                                   # [k]
        a2 = 0.5 * (a1 ** 2 - (x ** 2) @ (P ** 2))
        yp = x @ w + lams @ a2
        ls += L.sp_oracle_loss(loss_id, yp, y[i])
        dL = L.sp_oracle_dloss(loss_id, yp, y[i])
        gw += dL * x
        gP += (dL * lams)[None, :] * (x[:, None] * (a1[None, :] - P * x[:, None]))
    return gP, gw, ls


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from sparsepoly_b200.distributed import global_sum, local_batches
    rng = np.random.RandomState(0)
    n_loc, d, k, B = 23, 7, 3, 8
    Xs = [rng.randn(n_loc, d) * (rng.rand(n_loc, d) < 0.6) for _ in range(world)]
    ys = [np.sign(rng.randn(n_loc)) for _ in range(world)]
    X, y = Xs[rank], ys[rank]
    P = 0.01 * np.random.RandomState(1).randn(d, k)
    w = np.zeros(d)
    lams = np.ones(k)
    reg = O.Reg("squaredl12", d, k)
    alpha, beta, gamma, eta0 = 0.1, 0.1, 0.05, 0.2
    n_glob, = global_sum([n_loc])
    assert n_glob == world * n_loc
    it, losses = 1, []
    for epoch in range(2):
        ls_epoch = 0.0
        for b0, b1, b_glob in local_batches(n_loc, B, world):
            gP, gw, ls = _grad_deg2(X, y, range(b0, b1), P, w, lams, 1, O)
            tP, tw = torch.from_numpy(gP), torch.from_numpy(gw)
            dist.all_reduce(tP)
            dist.all_reduce(tw)
            ls_epoch += ls
            import ctypes as C
            a, b = C.c_double(), C.c_double()
            O.lib().sp_oracle_get_eta(1, eta0, alpha, beta, 1.0, it, C.byref(a), C.byref(b))
            eta_P, eta_w = a.value, b.value
            w = (w - gw * (eta_w / b_glob)) / (1 + eta_w * alpha)
            P = np.ascontiguousarray((P - gP * (eta_P / b_glob)) / (1.0 + eta_P * beta))
            reg.prox(P, gamma * eta_P / (1 + eta_P * beta))
            it += 1
        losses.append(global_sum([ls_epoch])[0] / n_glob)
    if rank == 0:
        np.savez(out, P=P, w=w, it=it, losses=np.array(losses),
                 X0=Xs[0], X1=Xs[1], y0=ys[0], y1=ys[1])
    dist.destroy_process_group()


def test_sharded_psgd_equals_single_process_oracle(tmp_path):
    from oracle import oracle as O
    from sparsepoly_b200.distributed import interleave_shards
    out = str(tmp_path / "res.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    n_loc, B = 23, 8
    order = interleave_shards([np.arange(n_loc), np.arange(n_loc, 2 * n_loc)], B)
    Xall = np.vstack([z["X0"], z["X1"]])[order]
    yall = np.concatenate([z["y0"], z["y1"]])[order]
    d, k = Xall.shape[1], 3
    P0 = 0.01 * np.random.RandomState(1).randn(d, k)
    ref = O.fit_fm(Xall, yall, degree=2, loss="logistic", n_components=k, solver="psgd",
                   regularizer="squaredl12", alpha=0.1, beta=0.1, gamma=0.05, fit_lower=None,
                   fit_linear=True, max_iter=2, tol=-1.0, batch_size=2 * (B // 2), eta0=0.2,
                   learning_rate="optimal", n_iter_no_change=10 ** 9, random_state=0,
                   P_init=np.ascontiguousarray(P0.T)[None])
    got_P = z["P"].T[None]
    assert np.max(np.abs(got_P - ref["P_"])) <= 1e-11 * max(np.max(np.abs(ref["P_"])), 1e-12) + 1e-15
    assert np.array_equal(got_P != 0, ref["P_"] != 0)
    assert np.max(np.abs(z["w"] - ref["w_"])) <= 1e-11
    assert int(z["it"]) == ref["it_"]
    assert np.allclose(z["losses"], ref["trace"], rtol=1e-10)


# ------------------------------------------------------------------ planned path: host logic over real collectives
def _plan_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparsepoly_b200 import distributed, synth
    from sparsepoly_b200.psgd_plan import PsgdPlan
    assert distributed.active_group() is None                     # initialised torch.distributed alone shards nothing
    group = distributed.enable_sharding()
    assert distributed.active_group() is group
    n_loc, d, b_loc = 300, 120, 64
    X = synth.criteo_like(n_loc, d, 50 + rank)
    csr = (torch.from_numpy(X.indptr.astype(np.int32)), torch.from_numpy(X.indices.astype(np.int32)),
           torch.from_numpy(X.data.astype(np.float64)))
    idx = torch.arange(n_loc, dtype=torch.int32)
    plan = PsgdPlan(csr, idx, d, b_loc, world=world, rank=rank, group=group)      # owner tables via all_gather (gloo)
    np.savez(os.path.join(out_dir, f"plan{rank}.npz"), own_q=plan.own_q.numpy(), own_src=plan.own_src.numpy(),
             mb_optr=plan.mb_optr, inbox_cap=plan.inbox_cap, mb_owner_start=plan.mb_owner_start,
             csr_slot=plan.csr_slot.numpy())
    # replicas start from rank 0's parameters whatever each rank drew
    P = np.random.RandomState(rank).randn(2, 3)
    it = np.array([float(7 + rank)])
    distributed.broadcast_arrays([P, it], group)
    assert np.array_equal(P, np.random.RandomState(0).randn(2, 3)) and it[0] == 7.0
    # unequal shards are refused
    if rank == 1:
        idx2 = torch.arange(n_loc - 1, dtype=torch.int32)
    else:
        idx2 = idx
    try:
        PsgdPlan(csr, idx2, d, b_loc, world=world, rank=rank, group=group)
        raise AssertionError("unequal shards accepted")
    except ValueError as e:
        assert "equal shards" in str(e)

    class _Est:                                                    # sharded_predict: rank-order concatenation, ragged shards
        def _predict(self, Xl):
            return np.asarray(Xl).sum(1)
    Xl = np.full((3 + rank, 2), float(rank + 1))
    got = distributed.sharded_predict(_Est(), Xl)
    want = np.concatenate([np.full(3 + r, 2.0 * (r + 1)) for r in range(world)])
    assert np.array_equal(got, want)
    assert np.array_equal(distributed.sharded_predict(_Est(), Xl, gather=False), np.full(3 + rank, 2.0 * (rank + 1)))
    distributed.disable_sharding()
    dist.destroy_process_group()


def test_plan_owner_tables_over_gloo_match_in_process_build(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from sparsepoly_b200 import synth
    from sparsepoly_b200.psgd_plan import PsgdPlan
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_plan_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    n_loc, d, b_loc = 300, 120, 64
    plans = []
    for r in range(world):
        X = synth.criteo_like(n_loc, d, 50 + r)
        csr = (torch.from_numpy(X.indptr.astype(np.int32)), torch.from_numpy(X.indices.astype(np.int32)),
               torch.from_numpy(X.data.astype(np.float64)))
        plans.append(PsgdPlan(csr, torch.arange(n_loc, dtype=torch.int32), d, b_loc, world=world, rank=r,
                              defer_owner_tables=True))
    lists = [p.column_lists() for p in plans]
    gathered = tuple([l[t] for l in lists] for t in range(4))
    for r, p in enumerate(plans):
        p.finish_owner_tables(*gathered)
        z = np.load(tmp_path / f"plan{r}.npz")
        assert np.array_equal(z["own_q"], p.own_q.numpy()) and np.array_equal(z["own_src"], p.own_src.numpy())
        assert np.array_equal(z["mb_optr"], p.mb_optr) and int(z["inbox_cap"]) == p.inbox_cap
        assert np.array_equal(z["mb_owner_start"], p.mb_owner_start) and np.array_equal(z["csr_slot"], p.csr_slot.numpy())
        # every row an owner lists is touched by at least one rank, and each (rank, index) appears once
        assert np.all((p.own_src.numpy() >= 0).any(1))
