"""Quick timing probe of the planned psgd path at the C5 shape (d=1M, k=32, 39 nnz/row): per-kernel-class
CUDA-event times over a few epochs.  python scripts/psgd_probe.py [rows] [epochs] [reg]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparsepoly_b200 import _lib, solvers, synth  # noqa: E402
from sparsepoly_b200.dataset import DeviceDataset  # noqa: E402
from sparsepoly_b200.psgd_plan import PsgdContext, PsgdPlan  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reg = sys.argv[3] if len(sys.argv) > 3 else "squaredl12"
gamma = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-6
d, k = 1_000_000, 32
t0 = time.time()
X = synth.criteo_like(rows, d, 4000)
y = (synth.planted_fm_targets(X, 4, 99, positive_frac=0.25) if os.environ.get('PROBE_PLANTED') == '1'
     else np.where(np.random.RandomState(99).rand(rows) < 0.25, 1.0, -1.0))
print(f"data {time.time()-t0:.1f}s nnz={X.nnz}", flush=True)
dev = torch.device("cuda", 0)
lib = _lib.load()
t0 = time.time()
ds = DeviceDataset(X, need_csr=True, need_csc=False, device=dev, hot_features=False)
torch.cuda.synchronize()
print(f"h2d {time.time()-t0:.2f}s", flush=True)
batch = int(rows * d / X.nnz)
idx = torch.arange(rows, dtype=torch.int32, device=dev)
t0 = time.time()
plan = PsgdPlan(ds.csr, idx, d, batch)
torch.cuda.synchronize()
print(f"plan {time.time()-t0:.2f}s M={plan.n_minibatches} cols/mb={plan.n_cols/plan.n_minibatches:.0f} "
      f"single/mb={len(plan.sg_u)/plan.n_minibatches:.0f} short/mb={len(plan.sc_u)/plan.n_minibatches:.0f} chunks/mb={len(plan.lc_u)/plan.n_minibatches:.0f} "
      f"multi/mb={len(plan.ml_u)/plan.n_minibatches:.0f} bytes={plan.nbytes()/1e9:.2f}GB", flush=True)
lams = torch.ones(k, dtype=torch.float64, device=dev)
ctx = PsgdContext(plan, 1, k, 2, reg, "logistic", True, lams)
P = torch.from_numpy(0.01 * np.random.RandomState(0).randn(1, d, k)).to(dev)
w = torch.zeros(d, dtype=torch.float64, device=dev)
ctx.load_model(P, w)
solvers.psgd_planned_begin(ctx)
yd = torch.from_numpy(y).to(dev)
loss = torch.zeros(1, dtype=torch.float64, device=dev)
it = 1
for ep in range(epochs):
    lib.sp_profile_enable(1)
    torch.cuda.synchronize()
    t0 = time.time()
    loss.zero_()
    it = solvers.psgd_planned_run(ctx, ds, plan, yd, idx, 1e-7, 1e-7, gamma, 0.1, 1, 1.0, it)
    solvers.psgd_planned_end(ctx, rows, loss, False)
    torch.cuda.synchronize()
    dt = time.time() - t0
    ms = (C.c_double * 8)()
    cnt = (C.c_longlong * 8)()
    lib.sp_profile_collect(ms, cnt)
    lib.sp_profile_enable(0)
    st = ctx.work[: 0].numel()
    print(f"epoch {ep}: {dt*1e3:.1f} ms  {rows/dt/1e6:.1f} M samples/s  loss {loss.item()/rows:.5f}  per-minibatch us: "
          f"rows {ms[4]/plan.n_minibatches*1e3:.1f} cols {ms[5]/plan.n_minibatches*1e3:.1f} stats {ms[6]/plan.n_minibatches*1e3:.1f} solve {ms[7]/plan.n_minibatches*1e3:.1f}", flush=True)
# unprofiled epoch (no event overhead)
torch.cuda.synchronize()
t0 = time.time()
it = solvers.psgd_planned_run(ctx, ds, plan, yd, idx, 1e-7, 1e-7, gamma, 0.1, 1, 1.0, it)
solvers.psgd_planned_end(ctx, rows, loss, True)
torch.cuda.synchronize()
dt = time.time() - t0
ctx.store_model(P, w)
print(f"plain epoch: {dt*1e3:.1f} ms  {rows/dt/1e6:.1f} M samples/s  nonzero frac {float((P != 0).double().mean()):.4f}", solvers.psgd_planned_solver_stats(ctx))
