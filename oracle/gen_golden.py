"""Generate tests/golden/*.npz by running the UNMODIFIED reference (numba path).

Run in the build container only (it imports /root/reference, which does not exist on the
GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

Each case stores the inputs (X as CSR triplet or dense, y, estimator kwargs as JSON) and the
reference's fitted attributes.  tests/test_oracle_golden.py replays every case through the C
oracle (pinning it) and tests/test_parity_gpu.py replays them through the CUDA backend.
"""
import json
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sp

REF = os.environ.get("SPARSEPOLY_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import sparsepoly  # noqa: E402  (the reference)
from sparsepoly.kernels import poly_predict  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def make_X(seed, n, d, density, dense=False):
    rng = np.random.RandomState(seed)
    if dense:
        return rng.randn(n, d)
    M = sp.random(n, d, density=density, format="csr", random_state=rng,
                  data_rvs=rng.randn).astype(np.float64)
    M.sort_indices()
    return M


def make_y(X, seed, kernel, degree, clf):
    rng = np.random.RandomState(seed + 1000)
    d = X.shape[1]
    Pt = rng.randn(3, d) * (rng.rand(3, d) < 0.5)
    if kernel == "anova":
        y = poly_predict(X, Pt, np.ones(3), "anova", degree)
    else:
        y = poly_predict(X, 0.3 * Pt, np.ones(3), "all-subsets")
    y = y + 0.1 * rng.randn(X.shape[0])
    if clf:
        y = np.where(y > np.median(y), 1.0, -1.0)
    return y


def pack_X(X):
    if sp.issparse(X):
        X = X.tocsr()
        return {"X_indptr": X.indptr.astype(np.int32), "X_indices": X.indices.astype(np.int32),
                "X_data": X.data.astype(np.float64), "X_shape": np.array(X.shape)}
    return {"X_dense": np.asarray(X)}


def tune_gamma(model, clf, kw, X, y, P_init=None):
    """Pick gamma (with the already-pinned C oracle, fast) so that the fitted P_ is neither
    all-zero nor fully dense nor numerically degenerate; the golden outputs themselves
    always come from the reference below."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_sp_oracle_py", os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle.py"))
    O = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(O)
    best = None
    for g in (0.3, 0.1, 0.03, 0.01, 3e-3, 1e-3, 3e-4, 1e-4, 3e-5, 1e-5):
        k2 = dict(kw, gamma=g)
        k2.setdefault("loss", "squared_hinge" if clf else "squared")
        if not clf:
            k2["loss"] = "squared"
        with np.errstate(all="ignore"):
            if P_init is not None:
                k2["P_init"] = P_init
            out = (O.fit_fm if model == "fm" else O.fit_all_subsets)(X, y, **k2)
        P = out["P_"]
        if not np.all(np.isfinite(P)):
            continue
        big = np.abs(P) > 1e-6
        frac = big.mean()
        tiny = ((P != 0) & ~big).mean()
        score = abs(frac - 0.5) + 5 * tiny
        if best is None or score < best[0]:
            best = (score, g, frac, tiny)
    return best[1], best[2], best[3]


def run_case(name, model, clf, kw, seed=0, n=60, d=12, density=0.3, dense=False):
    degree = kw.get("degree", 2)
    P_init = None
    if degree >= 4:
        # 0.01*randn init has a vanishing degree-4 gradient: warm-start from a larger P_
        density = 0.6
        n_orders = degree - 1 if kw.get("fit_lower", "explicit") == "explicit" else 1
        P_init = 0.4 * np.random.RandomState(seed + 77).randn(n_orders, kw["n_components"], d)
    X = make_X(seed, n, d, density, dense)
    y = make_y(X, seed, "anova" if model == "fm" else "all", degree, clf)
    if P_init is not None:
        y = y / 4.0 if not clf else y
    g, frac, tiny = tune_gamma(model, clf, kw, X, y, P_init)
    kw = dict(kw, gamma=g)
    print(f"   gamma={g} frac>1e-6={frac:.2f} tiny={tiny:.2f}")
    if model == "fm":
        cls = (sparsepoly.SparseFactorizationMachineClassifier if clf
               else sparsepoly.SparseFactorizationMachineRegressor)
    else:
        cls = (sparsepoly.SparseAllSubsetsClassifier if clf
               else sparsepoly.SparseAllSubsetsRegressor)
    est = cls(**kw)
    if P_init is not None:
        est.warm_start = True
        est.P_ = P_init.copy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        est.fit(X, y)
    assert np.all(np.isfinite(est.P_)), name
    rec = {"name": name, "model": model, "clf": bool(clf), "kw": kw}
    arrays = pack_X(X)
    arrays["y"] = y
    if P_init is not None:
        arrays["P_init"] = P_init
    arrays["P_"] = est.P_
    arrays["lams_"] = est.lams_
    if hasattr(est, "w_"):
        arrays["w_"] = est.w_
    arrays["n_iter_"] = np.array(est.n_iter_)
    if hasattr(est, "it_"):
        arrays["it_"] = np.array(est.it_)
    Xte = make_X(seed + 7, 25, d, density, dense)
    arrays["Xte"] = Xte.toarray() if sp.issparse(Xte) else Xte
    arrays["pred_te"] = (est.decision_function(Xte) if clf else est.predict(Xte))
    return rec, arrays


def cases():
    cs = []
    base = dict(n_components=4, max_iter=4, tol=-1.0, random_state=0)
    # ---- pcd FM
    i = 0
    for degree in (2, 3, 4):
        for reg in ("l1", "omegati"):
            for loss in ("squared", "logistic", "squared_hinge"):
                clf = loss != "squared"
                kw = dict(base, degree=degree, solver="pcd", regularizer=reg, beta=0.1, gamma=0.01,
                          alpha=0.1, fit_lower="explicit" if degree == 3 else None,
                          fit_linear=(i % 2 == 0), mean=(i % 3 == 0))
                if kw["mean"]:
                    kw.update(beta=0.1 / 60, alpha=0.1 / 60)
                if clf:
                    kw["loss"] = loss
                cs.append((f"pcd_fm_d{degree}_{reg}_{loss}", "fm", clf, kw,
                           dict(seed=i, dense=(i % 4 == 3))))
                i += 1
    for loss in ("squared", "logistic"):
        clf = loss != "squared"
        kw = dict(base, degree=2, solver="pcd", regularizer="squaredl12", beta=0.05, gamma=0.02,
                  fit_linear=True, init_lambdas="random_signs")
        if clf:
            kw["loss"] = loss
        cs.append((f"pcd_fm_sql12_{loss}", "fm", clf, kw, dict(seed=40 + clf)))
    kw = dict(base, degree=3, solver="pcd", regularizer="omegati", beta=0.1, gamma=0.05,
              fit_lower="explicit", fit_linear=True, shuffle=True, loss="logistic")
    cs.append(("pcd_fm_d3_shuffle", "fm", True, kw, dict(seed=50)))
    kw = dict(base, degree=3, solver="pcd", regularizer="l1", beta=0.1, gamma=0.01,
              fit_lower="augment", fit_linear=True)
    cs.append(("pcd_fm_d3_augment", "fm", False, kw, dict(seed=51)))
    # ---- pcd all-subsets
    for reg in ("l1", "omegati"):
        for loss in ("squared", "squared_hinge", "logistic"):
            clf = loss != "squared"
            kw = dict(n_components=4, max_iter=4, tol=-1.0, random_state=0, solver="pcd",
                      regularizer=reg, beta=0.1, gamma=0.01, eta0=0.5)
            if clf:
                kw["loss"] = loss
            cs.append((f"pcd_all_{reg}_{loss}", "all", clf, kw, dict(seed=60 + len(cs))))
    # ---- pbcd FM
    i = 0
    for degree in (2, 3):
        for reg in ("l1", "l21", "omegacs"):
            for loss in ("squared", "logistic", "squared_hinge"):
                clf = loss != "squared"
                kw = dict(base, degree=degree, solver="pbcd", regularizer=reg, beta=0.1,
                          gamma=0.01, alpha=0.1, fit_lower="explicit" if degree == 3 else None,
                          fit_linear=(i % 2 == 0), mean=(i % 3 == 0))
                if kw["mean"]:
                    kw.update(beta=0.1 / 60, alpha=0.1 / 60)
                if clf:
                    kw["loss"] = loss
                cs.append((f"pbcd_fm_d{degree}_{reg}_{loss}", "fm", clf, kw,
                           dict(seed=100 + i, dense=(i % 4 == 3))))
                i += 1
    kw = dict(base, degree=2, solver="pbcd", regularizer="squaredl21", beta=0.05, gamma=0.01)
    cs.append(("pbcd_fm_sql21", "fm", False, kw, dict(seed=130)))
    kw = dict(base, degree=4, solver="pbcd", regularizer="omegacs", beta=0.1, gamma=0.05,
              fit_lower=None, shuffle=True)
    cs.append(("pbcd_fm_d4_omegacs_shuffle", "fm", False, kw, dict(seed=131)))
    # ---- pbcd all-subsets
    for reg in ("l1", "l21", "omegacs"):
        kw = dict(n_components=4, max_iter=4, tol=-1.0, random_state=0, solver="pbcd",
                  regularizer=reg, beta=0.1, gamma=0.01, eta0=0.5)
        cs.append((f"pbcd_all_{reg}", "all", False, kw, dict(seed=140 + len(cs))))
    # ---- psgd
    i = 0
    for reg in ("l1", "l21", "squaredl12", "squaredl21"):
        for lr in ("constant", "optimal", "pegasos", "invscaling"):
            loss = ("squared", "logistic", "squared_hinge")[i % 3]
            clf = loss != "squared"
            degree = 2 + (i % 2)
            kw = dict(n_components=4, max_iter=6, tol=-1.0, random_state=0, degree=degree,
                      solver="psgd", regularizer=reg, alpha=0.1, beta=0.1, gamma=0.02,
                      fit_lower="explicit" if degree == 3 else None, fit_linear=(i % 2 == 0),
                      learning_rate=lr, eta0=0.2, batch_size=("auto", 1, 7, 60)[i % 4],
                      n_iter_no_change=1000, shuffle=(i % 5 == 4))
            if lr == "pegasos":   # eta = 1/(beta*it): needs a large beta to stay finite
                kw.update(alpha=2.0, beta=2.0)
            if clf:
                kw["loss"] = loss
            cs.append((f"psgd_{reg}_{lr}", "fm", clf, kw, dict(seed=200 + i)))
            i += 1
    return cs


def main():
    os.makedirs(OUT, exist_ok=True)
    index = []
    blob = {}
    for name, model, clf, kw, opts in cases():
        rec, arrays = run_case(name, model, clf, kw, **opts)
        index.append(rec)
        for key, val in arrays.items():
            blob[f"{name}/{key}"] = val
        print("ok", name)
    np.savez_compressed(os.path.join(OUT, "reference_fits.npz"), **blob)
    with open(os.path.join(OUT, "reference_fits.json"), "w") as f:
        json.dump(index, f, indent=1)
    # prox vectors (regularizer/utils.py:26-70 through SquaredL12.prox / SquaredL21.prox / L21 / L1)
    from sparsepoly.regularizer import L1, L21, SquaredL12, SquaredL21
    rng = np.random.RandomState(5)
    prox = {}
    for t, strength in enumerate((1e-3, 1e-2, 0.1, 1.0)):
        P = rng.randn(50, 6) * (rng.rand(50, 6) < 0.7)
        prox[f"in_{t}"] = P
        prox[f"strength_{t}"] = np.array(strength)
        for nm, cls in (("l1", L1), ("l21", L21), ("squaredl12", SquaredL12),
                        ("squaredl21", SquaredL21)):
            r = cls()
            r.init_cache_psgd(2, 50, 6)
            Q = P.copy()
            r.prox(Q, strength, 2)
            prox[f"{nm}_{t}"] = Q
    np.savez_compressed(os.path.join(OUT, "reference_prox.npz"), **prox)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
