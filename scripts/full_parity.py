"""Full-size parity of a bench workload against the C oracle (test infrastructure) on the same inputs.
usage: python scripts/full_parity.py pcd 1.0 1   (workload, scale, epochs)"""
import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import sparsepoly_b200 as S
from oracle import oracle as O
name, scale, epochs = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
X, y = bench.make_problem(name, scale, 0)
kw = dict(bench.WORKLOADS[name]["kw"], max_iter=epochs)
warnings.simplefilter("ignore")
cls = S.SparseFactorizationMachineClassifier if bench.WORKLOADS[name]["clf"] else S.SparseFactorizationMachineRegressor
for sweep in sys.argv[4:] or ["auto"]:
    os.environ["SPARSEPOLY_B200_SWEEP"] = sweep
    t0 = time.perf_counter(); est = cls(**kw).fit(X, y); t_gpu = time.perf_counter() - t0
    if sweep == (sys.argv[4:] or ["auto"])[0]:
        t0 = time.perf_counter(); out = O.fit_fm(X, y, **dict(kw, loss=kw.get("loss", "squared"))); t_cpu = time.perf_counter() - t0
    den = max(np.max(np.abs(out["P_"])), 1e-300)
    errP = float(np.max(np.abs(est.P_ - out["P_"])) / den)
    errw = float(np.max(np.abs(est.w_ - out["w_"])) / max(np.max(np.abs(out["w_"])), 1e-300))
    sup = bool(np.array_equal(est.P_ != 0, out["P_"] != 0))
    sd = (est.P_ != 0) != (out["P_"] != 0)
    ndiff = int(np.sum(sd))
    ndust = int(np.sum(sd & (np.maximum(np.abs(est.P_), np.abs(out["P_"])) < 1e-12 * den)))
    pl = est._dev_state["plan"]
    from sparsepoly_b200.objective import objective
    ob = objective(est, X, y)
    y_pm1 = est.label_binarizer_.transform(y).ravel().astype(np.float64) if hasattr(est, "label_binarizer_") else y
    okw = {k: kw[k] for k in ("degree", "regularizer", "alpha", "beta", "gamma", "mean", "fit_lower", "fit_linear") if k in kw}
    oo = O.objective_fm(X, y_pm1, out["P_"], out["w_"], est.lams_, loss=kw.get("loss", "squared"), **okw)
    print("objective (device, CUDA fit)", ob, "\nobjective (oracle, oracle fit)", oo,
          f"\nrelative difference of the total {abs(ob['total'] - oo['total']) / abs(oo['total']):.3e}", flush=True)
    print(f"{name} scale {scale} epochs {epochs} sweep={sweep} mode={pl.mode} {getattr(pl.wplan, 'stats', None) if pl.mode=='window' else ''}: "
          f"rel err P_ {errP:.3e}, w_ {errw:.3e}; supports identical: {sup} ({ndiff} of {est.P_.size} differ, {ndust} of them dust < 1e-12 max|P|); "
          f"nonzero frac {float(np.mean(out['P_'] != 0)):.4f}; gpu fit {t_gpu:.2f} s, oracle {t_cpu:.1f} s (1 core)", flush=True)
