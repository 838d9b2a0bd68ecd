"""One-off: nonzero fraction of P_ after 2 epochs for a few gamma values (freeze
beta / gamma so that 10-90 % of P_ is exactly zero).  usage: python scripts/tune_gamma.py pcd 1.0 1e-8 3e-9"""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import sparsepoly_b200 as S  # noqa: E402

name, scale = sys.argv[1], float(sys.argv[2])
X, y = bench.make_problem(name, scale, 0)
warnings.simplefilter("ignore")
for g in sys.argv[3:]:
    kw = dict(bench.WORKLOADS[name]["kw"], max_iter=2, gamma=float(g))
    cls = (S.SparseAllSubsetsClassifier if name == "allsub" else
           S.SparseFactorizationMachineClassifier if bench.WORKLOADS[name]["clf"] else
           S.SparseFactorizationMachineRegressor)
    est = cls(**kw).fit(X, y)
    P = est.P_ if est.P_.ndim == 3 else est.P_[None]
    print(name, "gamma", g, "nonzero frac per order", [round(float(np.mean(P[o] != 0)), 4) for o in range(P.shape[0])],
          "max|P|", float(np.abs(P).max()), flush=True)
