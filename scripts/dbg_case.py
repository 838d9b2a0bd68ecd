import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from golden_util import load_case, rel_err, same_support
import sparsepoly_b200 as S
name = sys.argv[1]
rec, X, arr = load_case(name)
cls = (S.SparseFactorizationMachineClassifier if rec["clf"] else S.SparseFactorizationMachineRegressor) if rec["model"] == "fm" else (S.SparseAllSubsetsClassifier if rec["clf"] else S.SparseAllSubsetsRegressor)
warnings.simplefilter("ignore")
for rep in range(3):
    est = cls(**rec["kw"])
    if arr.get("P_init") is not None:
        est.warm_start = True; est.P_ = arr["P_init"].copy()
    est.fit(X, arr["y"])
    pl = est._dev_state["plan"]
    print(name, os.environ.get("SPARSEPOLY_B200_WINDOW"), os.environ.get("SPARSEPOLY_B200_HORIZON"), os.environ.get("SPARSEPOLY_B200_NEAR"),
          pl.mode, getattr(pl.wplan, "stats", None), "relerr P", rel_err(est.P_, arr["P_"]), "support", same_support(est.P_, arr["P_"]),
          "w", rel_err(est.w_, arr["w_"]) if "w_" in arr else None, "X", X.shape, rec["kw"])
