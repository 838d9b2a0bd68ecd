import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import sparsepoly_b200 as S
from oracle import oracle as O
scale = float(sys.argv[1]); gam = float(sys.argv[2])
X, y = bench.make_problem("pcd", scale, 0)
warnings.simplefilter("ignore")
for epochs in (1, 2):
    kw = dict(bench.WORKLOADS["pcd"]["kw"], max_iter=epochs, gamma=gam)
    est = S.SparseFactorizationMachineClassifier(**kw).fit(X, y)
    out = O.fit_fm(X, y, **kw)
    for o in range(2):
        a, b = est.P_[o], out["P_"][o]
        diff = np.abs(a - b)
        print("epochs", epochs, "order", o, "max|P|", float(np.abs(b).max()), "max abs diff", float(diff.max()),
              "n diff>1e-9", int((diff > 1e-9).sum()), "support diffs", int(((a != 0) != (b != 0)).sum()),
              "dust-only support diffs", int((((a != 0) != (b != 0)) & (np.maximum(np.abs(a), np.abs(b)) < 1e-12)).sum()))
    print("  w max diff", float(np.abs(est.w_ - out["w_"]).max()), "pred range", float(np.abs(est.decision_function(X[:1000])).max()))
