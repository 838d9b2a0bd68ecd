"""Run-to-run reproducibility of the window sweep: two fits of the same problem must agree bit for bit
(fixed reduction trees, plan-determined summation order)."""
import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import sparsepoly_b200 as S
warnings.simplefilter("ignore")
for name, scale in (("pcd", 0.2), ("allsub", 0.3)):
    X, y = bench.make_problem(name, scale, 0)
    kw = dict(bench.WORKLOADS[name]["kw"], max_iter=2)
    cls = S.SparseAllSubsetsClassifier if name == "allsub" else S.SparseFactorizationMachineClassifier
    outs = []
    for rep in range(3):
        est = cls(**kw).fit(X, y)
        outs.append((est.P_.copy(), getattr(est, "w_", np.zeros(1)).copy()))
    pl = est._dev_state["plan"]
    same = all(np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1]) for o in outs[1:])
    print(name, scale, pl.mode, getattr(pl.wplan, "stats", None), "bitwise reproducible:", same,
          "nonzero frac", float(np.mean(outs[0][0] != 0)))
