"""Debug helper: per-position timeline (clock64) of window 100 of the LAST window sweep of a pcd fit
(library built with SP_WPROF=1) -- with >= 2 epochs of the C2 bench config that is a top-degree sweep
running entirely under zero-update speculation.
usage: SP_WPROF=1 python -m sparsepoly_b200.build --force; python scripts/wtrace_spec.py [scale] [epochs]"""
import ctypes as C
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import sparsepoly_b200 as S  # noqa: E402
from sparsepoly_b200 import _lib  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
X, y = bench.make_problem("pcd", scale, 0)
kw = dict(bench.WORKLOADS["pcd"]["kw"], max_iter=epochs)
only_top = len(sys.argv) > 3 and sys.argv[3] == "top"
if only_top:                      # epochs of 16 top-degree sweeps only (all speculative after the first)
    kw.update(fit_lower=None, fit_linear=False)
lib = _lib.load()
warnings.simplefilter("ignore")
est = S.SparseFactorizationMachineClassifier(**kw)
ws = (C.c_ulonglong * 2)()
lib.sp_wspec_read(ws)
NAMES = ["eng_wait", "eng_stage", "eng_role", "eng_flush", "ch_wait", "ch_comp", "wk_load", "wk_dep", "wk_terms",
         "wk_red", "wk_reswait", "wk_wb", "bulk_waitb", "bulk_base", "bulk_waitw", "bulk_wb"]
buf = (C.c_ulonglong * 16)()
if only_top:
    import torch, time
    from sklearn.utils import check_random_state
    Xc, yc = est._check_X_y(X, y)
    rng = check_random_state(kw["random_state"])
    est.w_ = np.zeros(Xc.shape[1])
    est.P_ = 0.01 * rng.randn(1, est.n_components, Xc.shape[1])
    est.lams_ = np.ones(est.n_components)
    epoch, sync = est._pcd_setup(Xc, np.ascontiguousarray(yc, dtype=np.float64), rng, torch.device("cuda", 0))
    for e in range(epochs):
        lib.sp_wprof_read(buf)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        epoch()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        lib.sp_wprof_read(buf)
        sync()
        nzf = float(np.mean(est.P_ != 0))
        coords = Xc.shape[1] * est.n_components
        print(f"epoch {e}: {dt:.3f} s, {dt / coords * 1e6:.3f} us/coordinate, nonzero {nzf:.4f}; engine cycles per coordinate:",
              {n: round(int(v) / coords, 1) for n, v in zip(NAMES, buf) if int(v)})
else:
    est.fit(X, y)
lib.sp_wspec_read(ws)
plan = est._dev_state["plan"]
print("plan", plan.mode, plan.wplan.stats, "speculated", ws[0], "rejected", ws[1],
      "nonzero by order", [float(np.mean(est.P_[o] != 0)) for o in range(est.P_.shape[0])])
tr = (C.c_longlong * (256 * 8))()
lib.sp_wtrace_read(tr)
B = plan.wplan.stats["window"]
T = np.array(list(tr), dtype=np.int64).reshape(256, 8)[:B]
base = T[:, 0].min()
mv = np.zeros(B, dtype=bool)
print(" tl  start  terms_done  cell_out  ch_seen  ch_done  res_seen  wb_flag  allsum_done | ch_done-prev")
for tl in range(min(B, 96)):
    r = T[tl] - base
    print(f"{tl:3d} {r[0]:6d} {r[1]:10d} {r[2]:9d} {r[3]:8d} {r[4]:8d} {r[5]:9d} {r[6]:8d} {r[7]:11d} | "
          f"{int(T[tl, 4] - T[tl - 1, 4]) if tl else 0:6d}")
dch = np.diff(T[:, 4])
print("chain step (ch_done - previous ch_done): mean", dch.mean(), "median", np.median(dch), "p90", np.percentile(dch, 90))
print("worker: start->terms", (T[:, 1] - T[:, 0]).mean(), "terms->allsum", (T[:, 7] - T[:, 1]).mean(),
      "allsum->cell_out", (T[:, 2] - T[:, 7]).mean(), "cell_out->ch_seen", (T[:, 3] - T[:, 2]).mean(),
      "ch_seen->ch_done", (T[:, 4] - T[:, 3]).mean(), "ch_done->res_seen", (T[:, 5] - T[:, 4]).mean(),
      "res_seen->wb_flag", (T[:, 6] - T[:, 5]).mean())
print("window span (first start -> last wb_flag)", int(T[:, 6].max() - base), "cycles for", B, "positions")
