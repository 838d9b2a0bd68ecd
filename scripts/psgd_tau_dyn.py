"""How fast do the squared-l1,2 thresholds move between minibatches (C5 shape)?  Sizing data for a band buffer."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sparsepoly_b200 import solvers, _lib
from sparsepoly_b200.dataset import DeviceDataset
bench.ROWS_OVERRIDE = 1500000
X, y = bench.make_problem("psgd", 1.0, 0)
wl = bench.WORKLOADS["psgd"]; kw = wl["kw"]; k = wl["k"]
n, d = X.shape
dev = torch.device("cuda", 0)
ds = DeviceDataset(X, need_csr=True, need_csc=False, device=dev)
yd = torch.from_numpy(y).to(dev); idx = torch.arange(n, dtype=torch.int32, device=dev)
rng = np.random.RandomState(0)
P = torch.from_numpy(np.ascontiguousarray(0.01 * rng.randn(1, d, k))).to(dev)
w = torch.zeros(d, dtype=torch.float64, device=dev); lams = torch.ones(k, dtype=torch.float64, device=dev)
gP, gw = torch.zeros_like(P), torch.zeros_like(w); loss = torch.zeros(1, dtype=torch.float64, device=dev)
work = solvers.prox_work(d, k, dev); st = solvers.PsgdLazyState(P, kw["regularizer"])
b = int(n * d / X.nnz)
prev = None
for m in range(50):
    solvers.psgd_minibatch(ds, yd, P, w, lams, 2, kw["alpha"], kw["beta"], kw["gamma"], kw["regularizer"], kw["loss"],
                           gP, gw, idx, True, kw["eta0"], 1, kw["power_t"], m * b, (m + 1) * b, b, m + 1, loss, work, st)
    thr = st.thr.clone()
    if prev is not None and m % 5 == 0:
        rel = ((thr - prev).abs() / prev.abs().clamp_min(1e-300))
        v = P[0, :, 0].abs(); t0 = float(thr[0])
        band = [int(((v > t0 * (1 - dl)) & (v <= t0 * (1 + dl))).sum()) for dl in (0.001, 0.01, 0.05)]
        print(f"mb {m}: thr[0] {t0:.4e}  rel change per step: median {float(rel.median()):.3e} max {float(rel.max()):.3e};"
              f" active {int((v > t0).sum())}; elements within +-0.1%/1%/5% of tau: {band}", flush=True)
    prev = thr
    if m % 5 == 0:
        ncol = k
        tail = 2 * 148 * 8 * ncol + 2 * ncol + 64
        stt = st.work[tail:tail + 8].cpu().numpy()
        bn = st.work[tail + 8 + ncol: tail + 8 + ncol + (ncol + 3) // 2 + 1].view(torch.int32)[:ncol + 2].cpu().numpy()
        print("   state: calls %d band hits %d generic %d delta %.4f; band counts (this call) min/max %d/%d fallback flags %d %d" % (stt[1], stt[2], stt[3], stt[4], bn[:ncol].min(), bn[:ncol].max(), bn[ncol], bn[ncol + 1]), flush=True)
