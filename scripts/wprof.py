"""Debug helper: cycle accounting of the window sweep's engine roles (library built with SP_WPROF=1).
usage: python scripts/wprof.py [scale] [epochs]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import sparsepoly_b200 as S  # noqa: E402
from sparsepoly_b200 import _lib  # noqa: E402

NAMES = ["eng_wait", "eng_stage", "eng_role", "eng_flush", "ch_wait", "ch_comp", "wk_load", "wk_dep", "wk_terms",
         "wk_red", "wk_reswait", "wk_wb", "bulk_waitb", "bulk_base", "bulk_waitw", "bulk_wb"]
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
name = sys.argv[3] if len(sys.argv) > 3 else "pcd"
X, y = bench.make_problem(name, scale, 0)
kw = dict(bench.WORKLOADS[name]["kw"], max_iter=epochs)
cls = S.SparseFactorizationMachineClassifier if bench.WORKLOADS[name]["clf"] else S.SparseFactorizationMachineRegressor
if name == "allsub":
    cls = S.SparseAllSubsetsClassifier
lib = _lib.load()
buf = (C.c_ulonglong * 16)()
import torch
import warnings
warnings.simplefilter("ignore")
est = cls(**kw)
lib.sp_wprof_read(buf)
torch.cuda.synchronize()
t0 = time.perf_counter()
est.fit(X, y)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
lib.sp_wprof_read(buf)
plan = est._dev_state["plan"]
print("fit wall", round(dt, 3), "s; plan", plan.mode, getattr(plan.wplan, "stats", None))
d = X.shape[1]
coords = d * (1 + kw["n_components"] * (kw.get("degree", 2) - 1)) * epochs if name == "pcd" else d * kw["n_components"] * epochs
tot = {n: int(v) for n, v in zip(NAMES, buf)}
for n in NAMES:
    print(f"{n:12s} {tot[n]/1.965e9:9.4f} s   {tot[n]/max(coords,1):9.1f} cyc/coordinate")

tr = (C.c_longlong * (256 * 8))()
lib.sp_wtrace_read(tr)
T = np.array(list(tr), dtype=np.int64).reshape(256, 8)
B = plan.wplan.stats["window"]
T = T[:B, :8]
base = T[:, 0].min()
print("per-position timeline of window 100 (cycles since first worker start):")
print(" tl  start  deps_ok  cell_out  ch_seen  ch_done  res_seen  wb_flag |  ch_done-prev  | allsum_done-deps_ok")
for tl in range(min(B, 64)):
    r = T[tl] - base
    print(f"{tl:3d} {r[0]:6d} {r[1]:8d} {r[2]:9d} {r[3]:8d} {r[4]:8d} {r[5]:9d} {r[6]:8d} | {int(T[tl,4]-T[tl-1,4]) if tl else 0:6d} | {int(T[tl,7]-T[tl,1]):6d}")
import torch
hp = plan.wplan.ht_ptr.cpu().numpy(); hd = plan.wplan.h_dep.cpu().numpy()
t0 = 100 * B
for tl in range(min(B, 24)):
    deps = hd[hp[t0 + tl]:hp[t0 + tl + 1]]
    print(tl, "n_hot", len(deps), "deps", sorted(int(x) for x in deps if x >= 0))
