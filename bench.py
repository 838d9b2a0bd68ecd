#!/usr/bin/env python
"""bench.py -- headline benchmark of the sparsepoly B200 backend (contract in the task prompt).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload auto|pcd|pbcd|allsub|psgd] [--scale S]

Workloads (BASELINE.json configs / SURVEY.md 8d):
  pcd    C2  FM-Clf degree=3 pcd omegati logistic, n=1M d=100k 50 nnz/row k=16   [N=1 default]
  pbcd   C3  FM-Reg degree=2 pbcd omegacs, n=1M d=100k 50 nnz/row k=32
  allsub C4  AllSubsets-Clf pcd omegati squared_hinge, n=500k d=20k 20 nnz/row k=16
  psgd   C5  FM-Clf degree=2 psgd squaredl12 logistic, Criteo-shaped d=1M 39 nnz/row k=32,
             sample-sharded, 6.25M rows per GPU (= n=50M at 8 GPUs)               [N>1 default]
A "step" is one epoch (pcd / pbcd / allsub) or one minibatch of batch_size="auto" (psgd).
pcd / pbcd do not shard (sequential coordinate order): with N>1 they run N independent replicas.
With --workload auto (default) the headline line is the pcd workload (replicas for N>1) and the
"also" list carries the psgd numbers (sample-sharded over the N ranks, NCCL all-reduce) and, at
N=1, the pbcd epoch time.

`--impl reference` times the reference algorithm on the host CPU (the pinned C oracle port of the
numba path; the numba package itself cannot travel to the GPU box) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "pcd": dict(tag="C2", n=1_000_000, d=100_000, r=50, seed=1, k=16, degree=3, clf=True,
                kw=dict(degree=3, loss="logistic", n_components=16, solver="pcd", regularizer="omegati",
                        alpha=1e-6, beta=1e-6, gamma=5e-10, mean=True, fit_linear=True,
                        fit_lower="explicit", shuffle=False, random_state=0, tol=-1.0)),
    "pbcd": dict(tag="C3", n=1_000_000, d=100_000, r=50, seed=2, k=32, degree=2, clf=False,
                 kw=dict(degree=2, n_components=32, solver="pbcd", regularizer="omegacs", alpha=1e-6,
                         beta=1e-6, gamma=1e-8, mean=True, fit_linear=True, fit_lower="explicit",
                         shuffle=False, random_state=0, tol=-1.0)),
    "allsub": dict(tag="C4", n=500_000, d=20_000, r=20, seed=3, k=16, degree=-1, clf=True,
                   kw=dict(loss="squared_hinge", n_components=16, solver="pcd", regularizer="omegati",
                           # OmegaTI for all-subsets multiplies gamma by prod_j (1 + |p_sj|) ~ e^160 at
                           # d=20k (reference omegati.py:19-34): only an absurdly small gamma leaves ~10 % of P_
                           beta=1e-6, gamma=1e-100, mean=True, shuffle=False, random_state=0, tol=-1.0)),
    "psgd": dict(tag="C5", n_per_gpu=6_250_000, d=1_000_000, r=39, seed=4, k=32, degree=2, clf=True,
                 kw=dict(degree=2, loss="logistic", n_components=32, solver="psgd",
                         regularizer="squaredl12", alpha=1e-7, beta=1e-7, gamma=1e-6, fit_linear=True,
                         fit_lower="explicit", batch_size="auto", eta0=0.1, learning_rate="optimal",
                         power_t=1.0, shuffle=False, random_state=0)),
}


# --------------------------------------------------------------------------------- utilities
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        out = {"sm_mhz": float(np.median(sm)) if sm else None,
               "sm_max_mhz": float(max(smax)) if smax else None, "reasons": sorted(reasons),
               "samples": len(sm)}
        if not sm:
            out["note"] = "timed region shorter than the 200 ms sampling period of nvidia-smi"
        return out


def profiled_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed `ncu --set full` capture (profiles/traffic.json); None when there is no capture."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel_key, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


ROWS_OVERRIDE = 0


def make_problem(name, scale=1.0, rank=0):
    """Synthetic inputs of the named config (host numpy / scipy)."""
    from sparsepoly_b200 import synth
    wl = WORKLOADS[name]
    if name == "psgd":
        n = max(1024, int((ROWS_OVERRIDE or wl["n_per_gpu"]) * scale))
        d = max(64, int(wl["d"] * scale))
        X = synth.criteo_like(n, d, wl["seed"] * 1000 + rank)
        rng = np.random.RandomState(99 + rank)
        y = np.where(rng.rand(n) < 0.25, 1.0, -1.0)       # CTR-like class balance
        return X, y
    n, d = max(256, int(wl["n"] * scale)), max(32, int(wl["d"] * scale))
    X = synth.uniform_sparse(n, d, wl["r"], wl["seed"] + 17 * rank)
    rng = np.random.RandomState(wl["seed"] + 100 + rank)
    # planted sparse linear + pairwise signal on 10% of the features (cheap host-side proxy for the
    # planted-model targets of SURVEY.md 8d; the timing does not depend on y)
    active = rng.choice(d, size=max(2, d // 10), replace=False)
    beta = np.zeros(d); beta[active] = rng.randn(active.size)
    s = X @ beta
    s = s + 0.5 * (s ** 2 - np.mean(s ** 2)) + 0.1 * np.std(s) * rng.randn(n)
    y = np.where(s > np.median(s), 1.0, -1.0) if wl["clf"] else (s / np.std(s))
    return X, y


def sweep_bytes(name, nnz, n, k, degree, fit_linear):
    """Algorithmic bytes of the SWEEP kernels in one epoch (SURVEY.md 8d byte model: column
    idx+val 12 B, A^1..A^(m-1) read+write 16(m-1) B, y_pred r/w 16 B, y 8 B per nonzero)."""
    lin = nnz * 36.0 if fit_linear else 0.0
    if name == "pcd":
        tot = lin
        for deg in range(2, degree + 1):
            tot += k * nnz * (12 + 16 * (deg - 1) + 24)
        return tot
    if name == "allsub":
        return k * nnz * (12 + 16 + 24)
    if name == "pbcd":
        return lin + nnz * (12 + 16 * k * (degree - 1) + 24)
    raise ValueError(name)


# --------------------------------------------------------------------------------- CPU arm
def cpu_reference_epoch_seconds(name, X, y, budget_s):
    """Time the reference algorithm (C oracle port, 1 host core -- the reference is single
    threaded) on a bounded sample and extrapolate to one full epoch of the workload.
    Sample = a leading slice of the columns (all rows), one component per order; CPU cost is
    linear in nnz x components.  Returns (seconds_per_epoch, description)."""
    from oracle import oracle as O
    import scipy.sparse as sp
    wl = WORKLOADS[name]
    kw = wl["kw"]
    n, d = X.shape
    k = wl["k"]
    # ~100 ns per nonzero-component-pass on one core: pick the column fraction for the budget
    passes = {"pcd": 1 + 2 * 5, "pbcd": 1 + 3 * 32 / 2.0, "allsub": 4}[name]
    frac = min(1.0, budget_s / (X.nnz * passes * 60e-9))
    d_s = max(16, int(d * frac))
    Xs = sp.csc_matrix(X[:, :d_s])
    csc = O.to_csc(Xs)
    frac = Xs.nnz / X.nnz
    rng = np.random.RandomState(0)
    idx_feat = np.arange(d_s, dtype=np.int32)
    a_, b_, g_ = (kw.get("alpha", 1) * n, kw["beta"] * n, kw["gamma"] * n)
    loss = kw.get("loss", "squared")
    y_pred = np.zeros(n)
    t_total = 0.0
    if name in ("pcd", "allsub"):
        reg = O.Reg(kw["regularizer"], d_s, 1)
        lams = np.ones(1)
        idx_comp = np.zeros(1, dtype=np.int32)
        if name == "pcd":
            w = np.zeros(d_s)
            cns = O.col_norm_sq(Xs)
            t0 = time.perf_counter()
            O.cd_linear_epoch(w, csc, y, y_pred, cns, a_, loss, idx_feat)
            t_lin = time.perf_counter() - t0
            t_comp = 0.0
            reg.init_pcd(wl["degree"])
            for deg in range(2, wl["degree"] + 1):
                P = 0.01 * rng.randn(1, d_s)
                A = np.zeros((n, deg + 1))
                t0 = time.perf_counter()
                O.pcd_epoch(P, csc, y, y_pred, lams, deg, b_, g_, 1.0, reg, loss, A, idx_comp, idx_feat)
                t_comp += time.perf_counter() - t0
            t_total = (t_lin + k * t_comp) / frac
        else:
            reg.init_pcd(-1)
            P = 0.01 * rng.randn(1, d_s)
            A = np.ones(n)
            y_pred = np.ones(n)
            t0 = time.perf_counter()
            O.pcd_all_epoch(P, csc, y, y_pred, lams, b_, g_, kw.get("eta0", 0.1), reg, loss, A, idx_comp, idx_feat)
            t_total = k * (time.perf_counter() - t0) / frac
        desc = (f"C oracle port of the numba path, 1 core: {d_s}/{d} leading columns ({frac:.3%} of nnz), "
                f"all {n} rows, 1 of {k} components per order, extrapolated linearly")
    else:  # pbcd
        kk = k
        reg = O.Reg(kw["regularizer"], d_s, kk)
        reg.init_pbcd(wl["degree"])
        w = np.zeros(d_s)
        cns = O.col_norm_sq(Xs)
        t0 = time.perf_counter()
        O.cd_linear_epoch(w, csc, y, y_pred, cns, a_, loss, idx_feat)
        P = np.ascontiguousarray(0.01 * rng.randn(d_s, kk))
        A = np.zeros((n, wl["degree"] + 1, kk))
        dA = np.zeros((n, wl["degree"], kk))
        O.pbcd_epoch(P, csc, y, y_pred, np.ones(kk), wl["degree"], b_, g_, 1.0, reg, loss, A, dA, idx_feat)
        t_total = (time.perf_counter() - t0) / frac
        desc = (f"C oracle port of the numba path, 1 core: {d_s}/{d} leading columns ({frac:.3%} of nnz), "
                f"all {n} rows, all {kk} components, extrapolated linearly")
    return t_total, desc


def cpu_reference_psgd_samples_per_s(X, y, budget_s):
    from oracle import oracle as O
    wl = WORKLOADS["psgd"]
    kw = wl["kw"]
    n, d = X.shape
    k = wl["k"]
    batch = int(n * d / X.nnz)
    # ~2 us/sample sparse part + dense update+prox ~ 3*d*k*8 B at ~5 GB/s effective per minibatch
    per_batch = batch * 3e-6 + d * k * 25e-9
    nb = max(1, int(budget_s / per_batch))
    ns = min(n, nb * batch)
    Xs = X[:ns]
    csr = O.to_csr(Xs)
    reg = O.Reg(kw["regularizer"], d, k)
    P = np.ascontiguousarray(0.01 * np.random.RandomState(0).randn(1, d, k))
    w = np.zeros(d)
    gP, gw = np.zeros_like(P), np.zeros(d)
    idx = np.arange(ns, dtype=np.int32)
    t0 = time.perf_counter()
    O.psgd_epoch(csr, y[:ns], P, w, np.ones(k), 2, kw["alpha"], kw["beta"], kw["gamma"], reg, kw["loss"],
                 gP, gw, idx, True, kw["eta0"], 1, kw["power_t"], batch, 1)
    dt = time.perf_counter() - t0
    return ns / dt, (f"C oracle port of the numba path, 1 core: first {ns} samples = {ns // batch} minibatches "
                     f"of {batch} at full d={d}, k={k}")


# --------------------------------------------------------------------------------- GPU arm
def run_sweep_workload(name, args, rank, world, local):
    import torch
    import sparsepoly_b200 as S
    from sparsepoly_b200 import _lib
    wl = WORKLOADS[name]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    X, y = make_problem(name, args.scale, rank)
    n, d = X.shape
    kw = dict(wl["kw"], max_iter=args.steps)
    if args.gamma is not None:
        kw["gamma"] = args.gamma
    if name == "allsub":
        cls = S.SparseAllSubsetsClassifier
    else:
        cls = S.SparseFactorizationMachineClassifier if wl["clf"] else S.SparseFactorizationMachineRegressor
    lib = _lib.load()

    # ---- device-resident timing: inputs in HBM, W warm-up + K timed epochs, CUDA events
    est = cls(**kw)
    Xc, yc = est._check_X_y(X, y)
    from sklearn.utils import check_random_state
    rng = check_random_state(kw["random_state"])
    lib.sp_set_device(local)
    if name == "allsub":
        est.P_ = 0.01 * rng.randn(est.n_components, Xc.shape[1])
        est.lams_ = np.ones(est.n_components)
        setup = est._setup
    else:
        Xc = est._augment(Xc)
        est.w_ = np.zeros(Xc.shape[1])
        n_orders = est.degree - 1 if est.fit_lower == "explicit" else 1
        est.P_ = 0.01 * rng.randn(n_orders, est.n_components, Xc.shape[1])
        est.lams_ = np.ones(est.n_components)
        setup = est._pcd_setup if name == "pcd" else est._pbcd_setup
    epoch, sync = setup(Xc, np.ascontiguousarray(yc, dtype=np.float64), rng, dev)
    for _ in range(args.warmup):
        epoch()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    lib.sp_profile_enable(1)
    import ctypes as C
    _z = (C.c_ulonglong * 2)()
    lib.sp_wspec_read(_z)                       # reset the speculation counters
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        epoch(read_back=False)
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    import ctypes as C
    ms = (C.c_double * 8)()
    cnt = (C.c_longlong * 8)()
    lib.sp_profile_collect(ms, cnt)
    lib.sp_profile_enable(0)
    wspec = (C.c_ulonglong * 2)()
    lib.sp_wspec_read(wspec)
    sync()
    nz_frac = float(np.mean(est.P_ != 0))
    nz_by_order = ([float(np.mean(est.P_[o] != 0)) for o in range(est.P_.shape[0])]
                   if est.P_.ndim == 3 else [nz_frac])
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    sec_per_epoch = ms_total / 1e3 / args.steps
    cls_id = 3 if name == "pbcd" else 2
    sweep_ms, sweep_n = float(ms[cls_id]) + (float(ms[2]) if name == "pbcd" else 0.0), int(cnt[cls_id])
    alg_bytes = sweep_bytes(name, X.nnz, n, wl["k"], wl["degree"], True) * args.steps
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (sweep_ms / 1e3) / 1e9 if sweep_ms > 0 else 0.0
    coords = (d * (1 + wl["k"] * (wl["degree"] - 1))) if name == "pcd" else (d * wl["k"] if name == "allsub" else 2 * d)
    result = {
        "value": sec_per_epoch, "ms_per_step": sec_per_epoch * 1e3,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "sweep_kernel (pcd.cu)" if name == "pcd" else "pbcd_sweep_kernel (pbcd.cu)",
                     "kernel_ms_per_step": sweep_ms / args.steps,
                     "kernel_share_of_step": sweep_ms / ms_total if ms_total else None,
                     "algorithmic_bytes_per_step": alg_bytes / args.steps,
                     "sequential_steps_per_epoch": coords,
                     "us_per_sequential_step": sweep_ms * 1e3 / args.steps / coords},
        "gpu_launches": int(sum(cnt)),
        "kernel_ms": {"rows": ms[0], "regcache": ms[1], "sweep_pcd": ms[2], "sweep_pbcd": ms[3]},
        "clocks": clocks, "p_nonzero_frac": nz_frac, "p_nonzero_frac_by_order": nz_by_order,
        "zero_update_speculation": {"positions": int(wspec[0]), "rejected": int(wspec[1]),
                                    "note": "window-sweep positions evaluated without per-record waits (pcd_window.cu)"},
        "geometry": ({"sweep": "window", **est._dev_state["plan"].wplan.stats}
                     if getattr(est._dev_state["plan"], "mode", "cluster") == "window" else
                     {"sweep": "cluster", "n_cta": est._dev_state["plan"].n_cta,
                      "threads": est._dev_state["plan"].threads}),
    }
    if result["geometry"]["sweep"] == "window" and name == "pbcd":
        result["roofline"]["kernel"] = "pbcd_wsweep_kernel (pbcd_window.cu)"
    if result["geometry"]["sweep"] == "window" and name in ("pcd", "allsub"):
        result["roofline"]["kernel"] = "wsweep_kernel (pcd_window.cu)"
        if name == "pcd" and args.scale == 1.0:
            result["roofline"]["traffic"] = profiled_traffic("wsweep_kernel")
        result["roofline"]["algorithmic_bytes_per_launch"] = alg_bytes / args.steps / max(1, sweep_n // args.steps)
    del est, epoch, sync
    torch.cuda.empty_cache()

    # ---- end to end through the public API: host buffers in, fitted host arrays out
    est2 = cls(**kw)
    import warnings
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        est2.fit(X, y)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    if rank == 0:
        # objective the fit reached (device evaluator, csrc/objective.cu; outside every timed region)
        from sparsepoly_b200.objective import objective
        result["objective"] = objective(est2, X, y)
    result["e2e"] = {"value": e2e_s, "unit": "s/epoch", "h2d_bytes_per_step": int(est2._h2d_bytes / args.steps),
                     "d2h_bytes_per_step": int((est2.P_.nbytes + getattr(est2, "w_", np.zeros(0)).nbytes) / args.steps + 8),
                     "note": f"fit(X_host, y_host) wall clock / {args.steps} epochs: host CSR->CSC, H2D, epochs, D2H"}
    if rank == 0 and not args.no_cpu:
        t_cpu, desc = cpu_reference_epoch_seconds(name, X, y, args.cpu_budget)
        result["cpu_baseline"] = {"value": t_cpu, "unit": "s/epoch", "cores": 1, "kind": "port", "sample": desc}
    result["config"] = {"workload": f"{wl['tag']} {name}: " + json.dumps({k: v for k, v in kw.items()}),
                        "n_samples": n, "n_features": d, "nnz": int(X.nnz), "scale": args.scale,
                        "l2": "inputs_exceed_l2 (CSC+CSR+records >> 126 MB)" if X.nnz * 24 > 2e8 else "inputs fit L2",
                        "parallelism": "single GPU" if world == 1 else f"{world} independent replicas (pcd/pbcd do not shard)"}
    return result


def run_psgd_workload(args, rank, world, local):
    import torch
    from sparsepoly_b200 import _lib, solvers
    from sparsepoly_b200.dataset import DeviceDataset
    wl = WORKLOADS["psgd"]
    kw = wl["kw"]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    lib.sp_set_device(local)
    group = None
    if world > 1:
        import torch.distributed as dist
        group = dist.group.WORLD
    global ROWS_OVERRIDE
    ROWS_OVERRIDE = args.rows_per_gpu
    X, y = make_problem("psgd", args.scale, rank)
    n, d = X.shape
    k = wl["k"]
    batch = int(n * d / X.nnz)                        # batch_size="auto" = d / nnz_row (independent of n)
    if args.psgd_batch == "auto":
        b_loc = max(1, batch // world)                # the reference's global minibatch, split over ranks
    elif args.psgd_batch == "weak":
        b_loc = batch                                 # every rank contributes one auto-sized batch
    else:
        b_loc = max(1, int(args.psgd_batch) // world)
    need = (args.warmup + args.steps) * b_loc
    if need > n:
        raise SystemExit(f"psgd bench needs {need} rows per GPU, shard has {n}")
    ds = DeviceDataset(X, need_csr=True, need_csc=False, device=dev)
    y_dev = torch.from_numpy(y).to(dev)
    idx = torch.arange(n, dtype=torch.int32, device=dev)
    rng = np.random.RandomState(0)
    P = torch.from_numpy(np.ascontiguousarray(0.01 * rng.randn(1, d, k))).to(dev)
    w = torch.zeros(d, dtype=torch.float64, device=dev)
    lams = torch.ones(k, dtype=torch.float64, device=dev)
    gP, gw = torch.zeros_like(P), torch.zeros_like(w)
    loss_dev = torch.zeros(1, dtype=torch.float64, device=dev)
    work = solvers.prox_work(d, k, dev)
    it = [1]
    state = solvers.PsgdLazyState(P, kw["regularizer"])

    def minibatch(m):
        b0, b1 = m * b_loc, (m + 1) * b_loc
        solvers.psgd_minibatch(ds, y_dev, P, w, lams, 2, kw["alpha"], kw["beta"], kw["gamma"],
                               kw["regularizer"], kw["loss"], gP, gw, idx, True, kw["eta0"], 1,
                               kw["power_t"], b0, b1, b_loc * world, it[0], loss_dev, work, state, group)
        it[0] += 1

    for m in range(args.warmup):
        minibatch(m)
    torch.cuda.synchronize()
    if group is not None:
        dist.barrier()
    lib.sp_profile_enable(1)
    import ctypes as C
    _z = (C.c_ulonglong * 2)()
    lib.sp_wspec_read(_z)                       # reset the speculation counters
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for m in range(args.warmup, args.warmup + args.steps):
        minibatch(m)
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    import ctypes as C
    ms = (C.c_double * 8)()
    cnt = (C.c_longlong * 8)()
    lib.sp_profile_collect(ms, cnt)
    lib.sp_profile_enable(0)
    if group is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    samples = args.steps * b_loc * world
    value = samples / (ms_total / 1e3)
    r = wl["r"]
    # SURVEY.md 8d byte model per sample / per minibatch
    per_sample = r * 12 + r * k * 8 + 2 * r * k * 8 + r * 8 * 3 + 8
    per_batch_dense = 5 * d * k * 8 + 4 * d * 8
    grad_bytes = per_sample * args.steps * b_loc
    peak, peak_src = measured_peak()
    grad_ms = float(ms[4])
    step_bytes = (per_sample * b_loc + per_batch_dense) * args.steps
    result = {
        "value": value, "ms_per_step": ms_total / args.steps,
        "roofline": {"bound": "hbm", "achieved": grad_bytes / (grad_ms / 1e3) / 1e9 if grad_ms else 0.0,
                     "peak": peak, "unit": "GB/s", "traffic": None, "peak_source": peak_src,
                     "kernel": "psgd_grad_kernel (psgd.cu)", "kernel_ms_per_step": grad_ms / args.steps,
                     "kernel_share_of_step": grad_ms / ms_total if ms_total else None,
                     "whole_step_gbs_per_gpu": step_bytes / (ms_total / 1e3) / 1e9,
                     "whole_step_frac": step_bytes / (ms_total / 1e3) / 1e9 / peak},
        "gpu_launches": int(sum(cnt)),
        "kernel_ms": {"psgd_grad": ms[4], "psgd_step_w": ms[5], "fused_update_prox": ms[6]},
        "clocks": clocks,
    }
    result["roofline"]["frac"] = result["roofline"]["achieved"] / peak
    # ---- e2e: pinned host CSR rows of each minibatch are copied inside the timed region
    Xr = X[: need]
    indptr_h = torch.from_numpy(Xr.indptr.astype(np.int32)).pin_memory()
    indices_h = torch.from_numpy(Xr.indices.astype(np.int32)).pin_memory()
    data_h = torch.from_numpy(Xr.data.astype(np.float64)).pin_memory()
    y_h = torch.from_numpy(y[:need].copy()).pin_memory()
    P.copy_(torch.from_numpy(np.ascontiguousarray(0.01 * rng.randn(1, d, k))))
    w.zero_(); it[0] = 1
    state = solvers.PsgdLazyState(P, kw["regularizer"])
    max_nnz = int(np.max(Xr.indptr[b_loc::b_loc] - Xr.indptr[:-b_loc:b_loc])) if need >= b_loc else Xr.nnz
    # two device buffer sets: minibatch m+1 is copied on a side stream while minibatch m is computed (the copy
    # of every step's inputs still happens inside the timed region, it just overlaps the previous step)
    bufs = [dict(ip=torch.empty(b_loc + 1, dtype=torch.int32, device=dev),
                 ix=torch.empty(max_nnz, dtype=torch.int32, device=dev),
                 dt=torch.empty(max_nnz, dtype=torch.float64, device=dev),
                 yb=torch.empty(b_loc, dtype=torch.float64, device=dev),
                 ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
    idx_b = torch.arange(b_loc, dtype=torch.int32, device=dev)
    copy_stream = torch.cuda.Stream(device=dev)
    n_mb = need // b_loc
    h2d = [0]

    def upload(m):
        """H2D of minibatch m's CSR rows and targets (pinned host memory) on the copy stream."""
        if m >= n_mb:
            return
        b = bufs[m & 1]
        r0, r1 = m * b_loc, (m + 1) * b_loc
        p0, p1 = int(indptr_h[r0]), int(indptr_h[r1])
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(b["free"])             # the compute that last read this buffer set
            b["ip"].copy_(indptr_h[r0:r1 + 1], non_blocking=True)
            b["ip"].sub_(p0)
            b["ix"][: p1 - p0].copy_(indices_h[p0:p1], non_blocking=True)
            b["dt"][: p1 - p0].copy_(data_h[p0:p1], non_blocking=True)
            b["yb"].copy_(y_h[r0:r1], non_blocking=True)
            b["ready"].record(copy_stream)
        h2d[0] += (b_loc + 1) * 4 + (p1 - p0) * 12 + b_loc * 8

    def minibatch_e2e(m, prefetch=True):
        b = bufs[m & 1]
        torch.cuda.current_stream().wait_event(b["ready"])
        if prefetch:
            upload(m + 1)                                 # overlaps this minibatch's kernels
        dsb = DeviceDataset.from_device_csr(b_loc, d, b["ip"], b["ix"], b["dt"])
        dsb.adopt_hot_features(ds)                        # dense-feature table of the training set
        solvers.psgd_minibatch(dsb, b["yb"], P, w, lams, 2, kw["alpha"], kw["beta"], kw["gamma"],
                               kw["regularizer"], kw["loss"], gP, gw, idx_b, True, kw["eta0"], 1,
                               kw["power_t"], 0, b_loc, b_loc * world, it[0], loss_dev, work, state, group)
        b["free"].record(torch.cuda.current_stream())
        it[0] += 1
        return loss_dev.item()                           # D2H read of the step's metric

    for b_ in bufs:
        b_["free"].record(torch.cuda.current_stream())
    upload(0)
    for m in range(args.warmup):
        minibatch_e2e(m, prefetch=m + 1 < args.warmup)    # (every timed step's copy happens inside the timed region)
    torch.cuda.synchronize()
    if group is not None:
        dist.barrier()
    h2d[0] = 0
    t0 = time.perf_counter()
    upload(args.warmup)
    for m in range(args.warmup, args.warmup + args.steps):
        minibatch_e2e(m)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if group is not None:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    result["e2e"] = {"value": samples / dt, "unit": "samples/s", "h2d_bytes_per_step": int(h2d[0] / args.steps),
                     "d2h_bytes_per_step": 8,
                     "note": "per minibatch: pinned host CSR rows + y -> device (double-buffered on a copy stream), gradient, (all-reduce), update, prox, loss read back"}
    if rank == 0 and not args.no_cpu:
        v, desc = cpu_reference_psgd_samples_per_s(X, y, args.cpu_budget)
        result["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": 1, "kind": "port", "sample": desc}
    result["config"] = {"workload": "C5 psgd: " + json.dumps(kw), "rows_per_gpu": n, "n_features": d,
                        "nnz_per_row": r, "global_batch": b_loc * world, "batch_mode": args.psgd_batch,
                        "batch_size_auto": batch, "scale": args.scale,
                        "l2": "inputs_exceed_l2 (P and grad_P are 256 MB each)" if d * k * 8 > 1.3e8 else "P fits L2",
                        "parallelism": "single GPU" if world == 1 else f"dp{world}: samples sharded, dense gradient all-reduced (NCCL) per minibatch"}
    return result


def run_reference(args, rank, world):
    if rank != 0:
        return None
    name = args.workload
    X, y = make_problem(name, args.scale, 0)
    per_step_budget = max(2.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    desc = ""
    t_run0 = time.perf_counter()
    for s in range(args.warmup + args.steps):
        if name == "psgd":
            v, desc = cpu_reference_psgd_samples_per_s(X, y, per_step_budget)
        else:
            v, desc = cpu_reference_epoch_seconds(name, X, y, per_step_budget)
        if s >= args.warmup:
            vals.append(v)
    wall = time.perf_counter() - t_run0
    value = float(np.mean(vals))
    unit = "samples/s" if name == "psgd" else "s/epoch"
    wl = WORKLOADS[name]
    return {"impl": "reference", "metric": metric_name(name), "value": value, "unit": unit,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall * 1e3 / max(1, args.steps + args.warmup),
            "higher_is_better": name == "psgd", "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{wl['tag']} {name}: " + json.dumps(wl["kw"]), "scale": args.scale},
            "cpu_baseline": {"value": value, "unit": unit, "cores": 1, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def metric_name(name):
    return {"pcd": "pcd_epoch_seconds", "pbcd": "pbcd_epoch_seconds", "allsub": "pcd_allsubsets_epoch_seconds",
            "psgd": "psgd_samples_per_second"}[name]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "pcd", "pbcd", "allsub", "psgd"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink n and d (debug only)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--rows-per-gpu", type=int, default=0, help="psgd: override the shard size (debug)")
    ap.add_argument("--psgd-batch", default="weak",
                    help="psgd global minibatch: 'auto' (= d/nnz_row split over the ranks), 'weak' "
                         "(auto x n_gpus: per-GPU work fixed) or an integer")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads")
    ap.add_argument("--gamma", type=float, default=None, help="override the workload's gamma (debug: other sparsity regimes)")
    args = ap.parse_args()
    rank, world, local = dist_env()
    global ROWS_OVERRIDE
    ROWS_OVERRIDE = args.rows_per_gpu
    auto = args.workload == "auto"
    if auto:
        args.workload = "pcd"
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out))
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = args.workload
    if name == "psgd":
        res = run_psgd_workload(args, rank, world, local)
    else:
        res = run_sweep_workload(name, args, rank, world, local)
    also = []
    if auto and not args.no_also:
        # secondary measurements of the same path (north_star: pbcd epoch time and psgd samples/s,
        # psgd sharded over the ranks); short runs, no CPU leg
        import copy
        a2 = copy.copy(args)
        a2.no_cpu = True
        a2.steps, a2.warmup = 20, 3
        if a2.rows_per_gpu == 0:
            a2.rows_per_gpu = 1_000_000          # >= 23 auto-sized minibatches; batch "auto" is n-independent
        for mode in (["auto"] if world == 1 else ["weak", "auto"]):
            a2.psgd_batch = mode
            r2 = run_psgd_workload(a2, rank, world, local)
            r2.update(metric=metric_name("psgd"), unit="samples/s", n_gpus=world)
            also.append(r2)
        if world == 1:
            a3 = copy.copy(args)
            a3.no_cpu = True
            a3.steps, a3.warmup = 2, 3
            r3 = run_sweep_workload("pbcd", a3, rank, world, local)
            r3.update(metric=metric_name("pbcd"), unit="s/epoch", n_gpus=1)
            also.append(r3)
    if rank == 0:
        line = {"metric": metric_name(name), "value": res.pop("value"),
                "unit": "samples/s" if name == "psgd" else "s/epoch", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res.pop("ms_per_step"),
                "higher_is_better": name == "psgd", "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic"}
        line.update(res)
        if also:
            line["also"] = also
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
