"""cProfile of the end-to-end fit() of the C5 psgd bench workload (host buffers in, host arrays out):
where the wall clock outside the epochs goes.  python scripts/fit_profile.py [rows] [epochs]"""
import cProfile
import os
import pstats
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import sparsepoly_b200 as S  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 6_250_000
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 20
bench.ROWS_OVERRIDE = rows
X, y = bench.make_problem("psgd", 1.0, 0)
kw = dict(bench.WORKLOADS["psgd"]["kw"], max_iter=epochs)
warnings.simplefilter("ignore")
S.SparseFactorizationMachineClassifier(**dict(kw, max_iter=1)).fit(X[:100000], y[:100000])     # context / library warm-up
t0 = time.perf_counter()
S.SparseFactorizationMachineClassifier(**kw).fit(X, y)
dt = time.perf_counter() - t0
print(f"plain fit: {dt:.3f} s for {epochs} epochs of {rows} rows = {rows * epochs / dt / 1e6:.1f} M samples/s", flush=True)
os.environ["SPARSEPOLY_B200_TIMING"] = "1"
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
est = S.SparseFactorizationMachineClassifier(**kw).fit(X, y)
pr.disable()
dt = time.perf_counter() - t0
print(f"fit: {dt:.3f} s for {epochs} epochs of {rows} rows = {rows * epochs / dt / 1e6:.1f} M samples/s", est._psgd_stats.get("setup_seconds"))
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
