"""Full-size parity of a bench workload against the C oracle (test infrastructure) on the same inputs.
usage: python scripts/full_parity.py <workload> <scale> <epochs> [rows]
  workload: pcd (C2) | pbcd (C3) | allsub (C4) | c1 (C1) | psgd (C5; `rows` = shard size, default 1M)
Prints relative errors of P_ / w_, support agreement, objective agreement and both wall clocks."""
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import sparsepoly_b200 as S  # noqa: E402
from oracle import oracle as O  # noqa: E402

name, scale, epochs = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
if name == "psgd":
    bench.ROWS_OVERRIDE = int(sys.argv[4]) if len(sys.argv) > 4 else 1_000_000
X, y = bench.make_problem(name, scale, 0)
wl = bench.WORKLOADS[name]
kw = dict(wl["kw"], max_iter=epochs)
warnings.simplefilter("ignore")
allsub = name == "allsub"
if allsub:
    cls = S.SparseAllSubsetsClassifier
else:
    cls = S.SparseFactorizationMachineClassifier if wl["clf"] else S.SparseFactorizationMachineRegressor
t0 = time.perf_counter()
est = cls(**kw).fit(X, y)
t_gpu = time.perf_counter() - t0
okw = dict(kw)
if not wl["clf"]:
    okw["loss"] = "squared"
y_pm1 = est.label_binarizer_.transform(y).ravel().astype(np.float64) if hasattr(est, "label_binarizer_") else y
t0 = time.perf_counter()
out = (O.fit_all_subsets if allsub else O.fit_fm)(X, y_pm1, **okw)
t_cpu = time.perf_counter() - t0
den = max(np.max(np.abs(out["P_"])), 1e-300)
errP = float(np.max(np.abs(est.P_ - out["P_"])) / den)
errw = float(np.max(np.abs(est.w_ - out["w_"])) / max(np.max(np.abs(out["w_"])), 1e-300)) if not allsub else 0.0
sd = (est.P_ != 0) != (out["P_"] != 0)
ndiff = int(np.sum(sd))
ndust = int(np.sum(sd & (np.maximum(np.abs(est.P_), np.abs(out["P_"])) < 1e-12 * den)))
from sparsepoly_b200.objective import objective  # noqa: E402
ob = objective(est, X, y)
if allsub:
    oo = O.objective_all_subsets(X, y_pm1, out["P_"], est.lams_, loss=kw["loss"], regularizer=kw["regularizer"], beta=kw["beta"],
                                 gamma=kw["gamma"], mean=kw.get("mean", False))
else:
    sel = {k: kw[k] for k in ("degree", "regularizer", "alpha", "beta", "gamma", "mean", "fit_lower", "fit_linear") if k in kw}
    if kw["solver"] == "psgd":
        sel["mean"] = False
    oo = O.objective_fm(X, y_pm1, out["P_"], out["w_"], est.lams_, loss=okw.get("loss", "squared"), **sel)
print("objective (device, CUDA fit)", ob, "\nobjective (oracle, oracle fit)", oo,
      f"\nrelative difference of the total {abs(ob['total'] - oo['total']) / abs(oo['total']):.3e}", flush=True)
extra = ""
if hasattr(est, "_dev_state"):
    pl = est._dev_state["plan"]
    extra = f" sweep mode={pl.mode}"
if hasattr(est, "it_"):
    extra += f" it_={est.it_} (oracle {out['it_']})"
print(f"{wl['tag']} {name} scale {scale} epochs {epochs} n={X.shape[0]} d={X.shape[1]} nnz={X.nnz}{extra}: "
      f"rel err P_ {errP:.3e}, w_ {errw:.3e}; supports identical: {ndiff == 0} ({ndiff} of {est.P_.size} differ, {ndust} of them "
      f"dust < 1e-12 max|P|); nonzero frac {float(np.mean(out['P_'] != 0)):.4f}; gpu fit {t_gpu:.2f} s, oracle {t_cpu:.1f} s (1 core)",
      flush=True)
