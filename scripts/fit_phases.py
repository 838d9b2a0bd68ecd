"""Debug helper: wall-clock phases of fit() at a bench workload (host conversion, H2D, plan, epochs)."""
import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import torch
import sparsepoly_b200 as S
from sparsepoly_b200 import dataset as D
name = sys.argv[1] if len(sys.argv) > 1 else "pcd"
X, y = bench.make_problem(name, 1.0, 0)
torch.cuda.synchronize()
def T(label, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
    print(f"{label:28s} {time.perf_counter()-t0:8.3f} s", flush=True); return r
T("warm import/lib", lambda: D._lib.load())
csr = T("host_csr", lambda: D.host_csr(X))
csc = T("host_csc (scipy tocsc)", lambda: D.host_csc(X))
dev = torch.device("cuda", 0)
T("h2d csr", lambda: [D._h2d(a, dev) for a in csr])
T("h2d csc", lambda: [D._h2d(a, dev) for a in csc])
ds = T("DeviceDataset", lambda: D.DeviceDataset(X, True, True, dev))
plan = T("SweepPlan ctor", lambda: D.SweepPlan(ds, "pcd", rec_stride=4))
T("plan.set_order", lambda: plan.set_order(np.arange(X.shape[1], dtype=np.int32)))
print(plan.mode, getattr(plan.wplan, "stats", None))
warnings.simplefilter("ignore")
kw = dict(bench.WORKLOADS[name]["kw"], max_iter=1)
cls = S.SparseFactorizationMachineClassifier if bench.WORKLOADS[name]["clf"] else S.SparseFactorizationMachineRegressor
T("fit (1 epoch)", lambda: cls(**kw).fit(X, y))
T("fit (1 epoch) again", lambda: cls(**kw).fit(X, y))
