"""Python face of the C parity oracle -- TEST INFRASTRUCTURE ONLY.

ctypes bindings for oracle/libsp_oracle.so plus host drivers that restate the reference's
fit loops (sparse_factorization_machines.py:94-353,355-451 and sparse_all_subsets.py:80-263)
so that a whole `fit` can be replayed on the CPU without the reference package (which does
not exist on the GPU box).  Parity status: pinned against the live reference through
tests/golden (see tests/test_oracle_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this module.  The product package sparsepoly_b200 never does.
"""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LOSS_IDS = {"squared": 0, "logistic": 1, "squared_hinge": 2}
REG_IDS = {"l1": 0, "l21": 1, "squaredl12": 2, "squaredl21": 3, "omegati": 4, "omegacs": 5}
LEARNING_RATE = {"constant": 0, "optimal": 1, "pegasos": 2, "invscaling": 3}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def _i(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_ip)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libsp_oracle.so")
        if not os.path.exists(path):
            import importlib.util
            spec = importlib.util.spec_from_file_location("_sp_oracle_build",
                                                          os.path.join(_HERE, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        L = C.CDLL(path)
        L.sp_oracle_reg_create.restype = C.c_void_p
        L.sp_oracle_reg_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.sp_oracle_reg_destroy.argtypes = [C.c_void_p]
        L.sp_oracle_reg_init_pcd.argtypes = [C.c_void_p, C.c_int]
        L.sp_oracle_reg_init_pbcd.argtypes = [C.c_void_p, C.c_int]
        L.sp_oracle_reg_prox.argtypes = [C.c_void_p, _dp, C.c_double]
        L.sp_oracle_loss.restype = C.c_double
        L.sp_oracle_loss.argtypes = [C.c_int, C.c_double, C.c_double]
        L.sp_oracle_dloss.restype = C.c_double
        L.sp_oracle_dloss.argtypes = [C.c_int, C.c_double, C.c_double]
        L.sp_oracle_cd_linear_epoch.restype = C.c_double
        L.sp_oracle_cd_linear_epoch.argtypes = [_dp, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp,
                                                C.c_double, C.c_int, _ip]
        L.sp_oracle_pcd_epoch.restype = C.c_double
        L.sp_oracle_pcd_epoch.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp,
                                          _dp, C.c_int, C.c_double, C.c_double, C.c_double,
                                          C.c_void_p, C.c_int, _dp, _ip, _ip]
        L.sp_oracle_pcd_all_epoch.restype = C.c_double
        L.sp_oracle_pcd_all_epoch.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp,
                                              _dp, _dp, C.c_double, C.c_double, C.c_double,
                                              C.c_void_p, C.c_int, _dp, _ip, _ip]
        L.sp_oracle_pbcd_epoch.restype = C.c_double
        L.sp_oracle_pbcd_epoch.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp,
                                           _dp, C.c_int, C.c_double, C.c_double, C.c_double,
                                           C.c_void_p, C.c_int, _dp, _dp, _ip]
        L.sp_oracle_pbcd_all_epoch.restype = C.c_double
        L.sp_oracle_pbcd_all_epoch.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp,
                                               _dp, _dp, C.c_double, C.c_double, C.c_double,
                                               C.c_void_p, C.c_int, _dp, _ip]
        L.sp_oracle_psgd_epoch.restype = C.c_double
        L.sp_oracle_psgd_epoch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp,
                                           _dp, _dp, _dp, C.c_int, C.c_double, C.c_double,
                                           C.c_double, C.c_void_p, C.c_int, _dp, _dp, _ip, C.c_int,
                                           C.c_double, C.c_int, C.c_double, C.c_int,
                                           C.POINTER(C.c_int64)]
        L.sp_oracle_kernel_rows.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, C.c_int, _dp]
        L.sp_oracle_get_eta.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_int64, _dp, _dp]
        L.sp_oracle_loss_sum.restype = C.c_double
        L.sp_oracle_loss_sum.argtypes = [C.c_int, C.c_int, _dp, _dp]
        L.sp_oracle_col_norm_sq.argtypes = [C.c_int, _ip, _dp, _dp]
        L.sp_oracle_branch_count.restype = C.c_longlong
        L.sp_oracle_branch_count.argtypes = [C.c_int]
        L.sp_oracle_reg_eval.restype = C.c_double
        L.sp_oracle_reg_eval.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        _LIB = L
    return _LIB


def branch_counts(reset=False):
    """How often the rounding-triggered recovery branches ran since the last reset: (omegacs negative
    dcache, omegacs.py:90-96; omegacs negative cache recompute, :75-76/:80-81; squaredl21 drift guard,
    squaredl21.py:49-50).  Used to make sure a fixture really enters them."""
    out = tuple(int(lib().sp_oracle_branch_count(t)) for t in range(3))
    if reset:
        lib().sp_oracle_branch_reset()
    return out


# --------------------------------------------------------------------------- data layouts
def _is_sparse(X):
    return sp.issparse(X)


def to_csc(X):
    """(indptr, indices, data) int32/int32/fp64 column-major, as dataset.py:119-134 builds it.
    Dense input stores every entry (dataset.py:39-57)."""
    if _is_sparse(X):
        Xc = sp.csc_matrix(X).astype(np.float64)
        Xc.sum_duplicates()
        Xc.sort_indices()
        return (np.ascontiguousarray(Xc.indptr, dtype=np.int32),
                np.ascontiguousarray(Xc.indices, dtype=np.int32),
                np.ascontiguousarray(Xc.data, dtype=np.float64))
    X = np.asarray(X, dtype=np.float64)
    n, d = X.shape
    indptr = (np.arange(d + 1, dtype=np.int64) * n).astype(np.int32)
    indices = np.tile(np.arange(n, dtype=np.int32), d)
    data = np.ascontiguousarray(X.T).ravel().copy()
    return indptr, indices, data


def to_csr(X):
    if _is_sparse(X):
        Xr = sp.csr_matrix(X).astype(np.float64)
        Xr.sum_duplicates()
        Xr.sort_indices()
        return (np.ascontiguousarray(Xr.indptr, dtype=np.int32),
                np.ascontiguousarray(Xr.indices, dtype=np.int32),
                np.ascontiguousarray(Xr.data, dtype=np.float64))
    X = np.asarray(X, dtype=np.float64)
    n, d = X.shape
    indptr = (np.arange(n + 1, dtype=np.int64) * d).astype(np.int32)
    indices = np.tile(np.arange(d, dtype=np.int32), n)
    data = np.ascontiguousarray(X).ravel().copy()
    return indptr, indices, data


# ------------------------------------------------------------- prediction (kernels.py:43-153)
def _safe_power(X, degree):
    if _is_sparse(X):
        return X.power(degree)
    return X ** degree


def _D(X, P, degree):
    r = _safe_power(X, degree) @ (P.T ** degree)
    return np.asarray(r)


def _homogeneous(X, P, degree):
    K = np.asarray(X @ P.T, dtype=np.float64)
    if _is_sparse(K):
        K = K.toarray()
    K = K * 1.0
    K += 0
    K **= degree
    return K


def anova_kernel(X, P, degree=2):
    """kernels.py:71-115 (power-sum / Newton-Girard form, as the reference computes it)."""
    if degree == 2:
        K = _homogeneous(X, P, 2)
        K -= _D(X, P, 2)
        K /= 2
    elif degree == 3:
        K = _homogeneous(X, P, 3)
        K -= 3 * _D(X, P, 2) * _D(X, P, 1)
        K += 2 * _D(X, P, 3)
        K /= 6
    else:
        n1, n2 = X.shape[0], P.shape[0]
        Ds = [np.asarray(X @ P.T)]
        Ds += [_D(X, P, t) for t in range(2, degree + 1)]
        anovas = [1.0, Ds[0]]
        for m in range(2, degree + 1):
            anova = np.zeros((n1, n2))
            sign = 1.0
            for t in range(1, m + 1):
                anova += sign * anovas[m - t] * Ds[t - 1]
                sign *= -1.0
            anova /= 1.0 * m
            anovas.append(anova)
        K = anovas[-1]
    return K


def kernel_rows_dp(X, P_dk, degree):
    """Per-(sample, component) kernel values by the reference's DP (psgd.py:34-44 /
    kernels.py:118-129); degree=-1 is all-subsets.  P_dk is [d,k]."""
    indptr, indices, data = to_csr(X)
    n = len(indptr) - 1
    P_dk = np.ascontiguousarray(P_dk, dtype=np.float64)
    k = P_dk.shape[1]
    out = np.empty((n, k))
    lib().sp_oracle_kernel_rows(n, k, _i(indptr), _i(indices), _d(data), _d(P_dk), degree, _d(out))
    return out


def all_subsets_kernel(X, P):
    return kernel_rows_dp(X, np.ascontiguousarray(P.T), -1)


def poly_predict(X, P, lams, kernel, degree=2):
    """kernels.py:140-153.  P is [k,d]."""
    if kernel == "anova":
        K = anova_kernel(X, P, degree)
    elif kernel == "all-subsets":
        K = all_subsets_kernel(X, P)
    else:
        raise ValueError("Unsuppported kernel: {}".format(kernel))
    return np.dot(K, lams)


def col_norm_sq(X):
    """row_norms(X.T, squared=True) (sparse_factorization_machines.py:409)."""
    if _is_sparse(X):
        # sklearn's csr_row_norms on X.T: one sequential sum per feature, samples ascending (NOT scipy's
        # pairwise reduction -- the order matters at the ulp level and pcd amplifies it at n >= 10^4)
        indptr, _, data = to_csc(X)
        out = np.zeros(X.shape[1])
        lib().sp_oracle_col_norm_sq(int(X.shape[1]), _i(indptr), _d(data), _d(out))
        return out
    X = np.asarray(X, dtype=np.float64)
    return np.einsum("ij,ij->j", X, X)


class Reg:
    """Owns one sp_reg (the jitclass instance created per fit, base.py:27-34)."""

    def __init__(self, name, n_features, n_components):
        self.name = name
        self.h = lib().sp_oracle_reg_create(REG_IDS[name], n_features, n_components)

    def init_pcd(self, degree):
        if lib().sp_oracle_reg_init_pcd(self.h, degree) != 0:
            raise ValueError(f"regularizer {self.name} unsupported for pcd degree={degree}")

    def init_pbcd(self, degree):
        if lib().sp_oracle_reg_init_pbcd(self.h, degree) != 0:
            raise ValueError(f"regularizer {self.name} unsupported for pbcd degree={degree}")

    def prox(self, P_dk, strength):
        assert P_dk.flags.c_contiguous
        lib().sp_oracle_reg_prox(self.h, _d(P_dk), float(strength))

    def __del__(self):
        try:
            lib().sp_oracle_reg_destroy(self.h)
        except Exception:
            pass


# --------------------------------------------------------------------------- epoch wrappers
def cd_linear_epoch(w, csc, y, y_pred, cns, alpha, loss, idx_feat):
    indptr, indices, data = csc
    return lib().sp_oracle_cd_linear_epoch(_d(w), len(indptr) - 1, _i(indptr), _i(indices), _d(data),
                                           _d(y), _d(y_pred), _d(cns), float(alpha),
                                           LOSS_IDS[loss], _i(idx_feat))


def pcd_epoch(P_kd, csc, y, y_pred, lams, degree, beta, gamma, eta, reg, loss, A, idx_comp, idx_feat):
    indptr, indices, data = csc
    k, d = P_kd.shape
    return lib().sp_oracle_pcd_epoch(_d(P_kd), len(y), d, k, _i(indptr), _i(indices), _d(data), _d(y),
                                     _d(y_pred), _d(lams), degree, float(beta), float(gamma),
                                     float(eta), reg.h, LOSS_IDS[loss], _d(A), _i(idx_comp),
                                     _i(idx_feat))


def pcd_all_epoch(P_kd, csc, y, y_pred, lams, beta, gamma, eta, reg, loss, A, idx_comp, idx_feat):
    indptr, indices, data = csc
    k, d = P_kd.shape
    return lib().sp_oracle_pcd_all_epoch(_d(P_kd), len(y), d, k, _i(indptr), _i(indices), _d(data),
                                         _d(y), _d(y_pred), _d(lams), float(beta), float(gamma),
                                         float(eta), reg.h, LOSS_IDS[loss], _d(A), _i(idx_comp),
                                         _i(idx_feat))


def pbcd_epoch(P_dk, csc, y, y_pred, lams, degree, beta, gamma, eta, reg, loss, A, dA, idx_feat):
    indptr, indices, data = csc
    d, k = P_dk.shape
    return lib().sp_oracle_pbcd_epoch(_d(P_dk), len(y), d, k, _i(indptr), _i(indices), _d(data),
                                      _d(y), _d(y_pred), _d(lams), degree, float(beta), float(gamma),
                                      float(eta), reg.h, LOSS_IDS[loss], _d(A), _d(dA), _i(idx_feat))


def pbcd_all_epoch(P_dk, csc, y, y_pred, lams, beta, gamma, eta, reg, loss, A, idx_feat):
    indptr, indices, data = csc
    d, k = P_dk.shape
    return lib().sp_oracle_pbcd_all_epoch(_d(P_dk), len(y), d, k, _i(indptr), _i(indices), _d(data),
                                          _d(y), _d(y_pred), _d(lams), float(beta), float(gamma),
                                          float(eta), reg.h, LOSS_IDS[loss], _d(A), _i(idx_feat))


def psgd_epoch(csr, y, P_odk, w, lams, degree, alpha, beta, gamma, reg, loss, grad_P, grad_w,
               idx_samples, fit_linear, eta0, learning_rate, power_t, batch_size, it):
    indptr, indices, data = csr
    n_orders, d, k = P_odk.shape
    it_c = C.c_int64(int(it))
    s = lib().sp_oracle_psgd_epoch(len(y), d, k, n_orders, _i(indptr), _i(indices), _d(data), _d(y),
                                   _d(P_odk), _d(w), _d(lams), degree, float(alpha), float(beta),
                                   float(gamma), reg.h, LOSS_IDS[loss], _d(grad_P), _d(grad_w),
                                   _i(idx_samples), int(bool(fit_linear)), float(eta0),
                                   int(learning_rate), float(power_t), int(batch_size),
                                   C.byref(it_c))
    return s, it_c.value


# ------------------------------------------------------------------------------ fit drivers
def _augment(X, fit_lower, fit_linear, degree):
    """sparse_factorization_machines.py:86-92."""
    if fit_lower == "augment":
        k = 2 if fit_linear else 1
        for _ in range(degree - k):
            n = X.shape[0]
            if _is_sparse(X):
                X = sp.hstack([sp.csr_matrix(np.ones((n, 1))), X]).tocsr()
            else:
                X = np.hstack([np.ones((n, 1)), X])
    return X


def fm_output(X, P_okd, w, lams, degree, fit_linear, fit_lower):
    """_get_output, sparse_factorization_machines.py:437-451."""
    y_pred = poly_predict(X, P_okd[0], lams, "anova", degree)
    if fit_linear:
        y_pred = y_pred + np.asarray(X @ w).ravel()
    if fit_lower == "explicit" and degree == 3:
        y_pred = y_pred + poly_predict(X, P_okd[1], lams, "anova", 2)
    return np.ascontiguousarray(y_pred, dtype=np.float64)


def fit_fm(X, y, degree=2, loss="squared", n_components=2, solver="pcd", regularizer="squaredl12",
           alpha=1, beta=1, gamma=1, mean=False, tol=1e-6, fit_lower="explicit", fit_linear=True,
           init_lambdas="ones", max_iter=100, shuffle=False, batch_size="auto", eta0=1.0,
           learning_rate="optimal", power_t=1.0, n_iter_no_change=5, random_state=None,
           P_init=None, w_init=None, lams_init=None, it_init=1):
    """Restates fit + _fit_pcd/_fit_pbcd/_fit_psgd (sparse_factorization_machines.py:94-435).
    y must already be the float64 target vector (+-1 for classifiers).  Returns a dict with
    P_ [n_orders,k,d], w_, lams_, n_iter_, it_, y_pred (pcd/pbcd), viols / losses per epoch."""
    from sklearn.utils import check_random_state
    y = np.ascontiguousarray(y, dtype=np.float64)
    X = _augment(X, fit_lower, fit_linear, degree)
    n, d = X.shape
    k = n_components
    rng = check_random_state(random_state)
    w_ = np.zeros(d) if w_init is None else np.array(w_init, dtype=np.float64)
    n_orders = degree - 1 if fit_lower == "explicit" else 1
    P_ = 0.01 * rng.randn(n_orders, k, d) if P_init is None else np.array(P_init, dtype=np.float64)
    if lams_init is not None:
        lams_ = np.array(lams_init, dtype=np.float64)
    elif init_lambdas == "ones":
        lams_ = np.ones(k)
    elif init_lambdas == "random_signs":
        lams_ = np.sign(rng.randn(k))
    else:
        raise ValueError("bad init_lambdas")
    reg = Reg(regularizer, d, k)
    out = {"lams_": lams_, "trace": []}
    converged = False
    it = 0
    if solver in ("pcd", "pbcd"):
        csc = to_csc(X)
        y_pred = fm_output(X, P_, w_, lams_, degree, fit_linear, fit_lower)
        cns = np.ascontiguousarray(col_norm_sq(X))
        a_, b_, g_ = (alpha * n, beta * n, gamma * n) if mean else (alpha, beta, gamma)
        idx_feat = np.arange(d, dtype=np.int32)
    if solver == "pcd":
        idx_comp = np.arange(k, dtype=np.int32)
        A = np.zeros((n, degree + 1))
        reg.init_pcd(degree)
        for it in range(max_iter):
            viol = 0
            if shuffle:
                rng.shuffle(idx_comp)
                rng.shuffle(idx_feat)
            if fit_linear:
                viol += cd_linear_epoch(w_, csc, y, y_pred, cns, a_, loss, idx_feat)
            if fit_lower == "explicit":
                for deg in range(2, degree):
                    viol += pcd_epoch(P_[degree - deg], csc, y, y_pred, lams_, deg, b_, g_, eta0,
                                      reg, loss, A, idx_comp, idx_feat)
            viol += pcd_epoch(P_[0], csc, y, y_pred, lams_, degree, b_, g_, eta0, reg, loss, A,
                              idx_comp, idx_feat)
            out["trace"].append(viol)
            if viol < tol:
                converged = True
                break
        out["y_pred"] = y_pred
    elif solver == "pbcd":
        A = np.zeros((n, degree + 1, k))
        dA = np.zeros((n, degree, k))
        reg.init_pbcd(degree)
        P = np.ascontiguousarray(P_.swapaxes(1, 2))
        for it in range(max_iter):
            viol = 0
            if shuffle:
                rng.shuffle(idx_feat)
            if fit_linear:
                viol += cd_linear_epoch(w_, csc, y, y_pred, cns, a_, loss, idx_feat)
            if fit_lower == "explicit":
                for deg in range(2, degree):
                    viol += pbcd_epoch(P[degree - deg], csc, y, y_pred, lams_, deg, b_, g_, eta0,
                                       reg, loss, A, dA, idx_feat)
            viol += pbcd_epoch(P[0], csc, y, y_pred, lams_, degree, b_, g_, eta0, reg, loss, A, dA,
                               idx_feat)
            out["trace"].append(viol)
            if viol < tol:
                converged = True
                break
        P_[:, :, :] = np.array(P.swapaxes(1, 2))
        out["y_pred"] = y_pred
    elif solver == "psgd":
        csr = to_csr(X)
        it_ = it_init
        idx_samples = np.arange(n, dtype=np.int32)
        nnz = len(csr[2])
        bs = int(n * d / nnz) if batch_size == "auto" else batch_size
        lr = LEARNING_RATE[learning_rate]
        P = np.ascontiguousarray(P_.swapaxes(1, 2))
        grad_P = np.zeros(P.shape)
        grad_w = np.zeros(d)
        no_improve, best = 0, np.inf
        for it in range(max_iter):
            if shuffle:
                rng.shuffle(idx_samples)
            sum_loss, it_ = psgd_epoch(csr, y, P, w_, lams_, degree, alpha, beta, gamma, reg, loss,
                                       grad_P, grad_w, idx_samples, fit_linear, eta0, lr, power_t,
                                       bs, it_)
            sum_loss /= n
            out["trace"].append(sum_loss)
            if sum_loss > (best - tol):
                no_improve += 1
            else:
                no_improve = 0
            if sum_loss < best:
                best = sum_loss
            if no_improve >= n_iter_no_change:
                converged = True
                break
        P_[:, :, :] = np.array(P.swapaxes(1, 2))
        out["it_"] = it_
    else:
        raise ValueError(f"Solver {solver} is not supported.")
    out.update(P_=P_, w_=w_, n_iter_=it, converged=converged)
    return out


def fit_all_subsets(X, y, loss="squared", n_components=2, solver="pcd", beta=1, gamma=1, eta0=0.1,
                    mean=False, tol=1e-6, regularizer="omegati", init_lambdas="ones", max_iter=100,
                    shuffle=False, random_state=None, P_init=None, lams_init=None):
    """Restates sparse_all_subsets.py:80-263."""
    from sklearn.utils import check_random_state
    y = np.ascontiguousarray(y, dtype=np.float64)
    n, d = X.shape
    k = n_components
    rng = check_random_state(random_state)
    P_ = 0.01 * rng.randn(k, d) if P_init is None else np.array(P_init, dtype=np.float64)
    if lams_init is not None:
        lams_ = np.array(lams_init, dtype=np.float64)
    elif init_lambdas == "ones":
        lams_ = np.ones(k)
    else:
        lams_ = np.sign(rng.randn(k))
    csc = to_csc(X)
    y_pred = np.ascontiguousarray(poly_predict(X, P_, lams_, "all-subsets"))
    b_, g_ = (beta * n, gamma * n) if mean else (beta, gamma)
    reg = Reg(regularizer, d, k)
    idx_feat = np.arange(d, dtype=np.int32)
    out = {"lams_": lams_, "trace": []}
    converged = False
    it = 0
    if solver == "pcd":
        idx_comp = np.arange(k, dtype=np.int32)
        A = np.ones(n)
        reg.init_pcd(-1)
        for it in range(max_iter):
            if shuffle:
                rng.shuffle(idx_comp)
                rng.shuffle(idx_feat)
            viol = pcd_all_epoch(P_, csc, y, y_pred, lams_, b_, g_, eta0, reg, loss, A, idx_comp,
                                 idx_feat)
            out["trace"].append(viol)
            if viol < tol:
                converged = True
                break
    elif solver == "pbcd":
        A = np.ones((n, k))
        reg.init_pbcd(-1)
        P = np.ascontiguousarray(P_.T)
        for it in range(max_iter):
            if shuffle:
                rng.shuffle(idx_feat)
            viol = pbcd_all_epoch(P, csc, y, y_pred, lams_, b_, g_, eta0, reg, loss, A, idx_feat)
            out["trace"].append(viol)
            if viol < tol:
                converged = True
                break
        P_[:, :] = np.array(P.T)
    else:
        raise ValueError(f"Solver {solver} is not supported.")
    out.update(P_=P_, y_pred=y_pred, n_iter_=it, converged=converged)
    return out


# --------------------------------------------------------------------------- objective
def loss_sum(y_pred, y, loss):
    """sum_i loss(y_pred_i, y_i) (loss.py:19-20, :34-41, :61-65), ascending i."""
    y_pred = np.ascontiguousarray(y_pred, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    return float(lib().sp_oracle_loss_sum(len(y), LOSS_IDS[loss], _d(y_pred), _d(y)))


def reg_eval(P_kd, regularizer, degree):
    """Omega(P) of one order given as the reference stores it, P_kd [k,d] (regularizer `eval`
    methods: l1.py:17-18, l21.py:19-21, squaredl12.py:20-22, squaredl21.py:23-25, omegati.py:19-47,
    omegacs.py:22-39); degree=-1 for the all-subsets model."""
    P_dk = np.ascontiguousarray(np.asarray(P_kd, dtype=np.float64).T)
    d, k = P_dk.shape
    return float(lib().sp_oracle_reg_eval(REG_IDS[regularizer], int(degree), d, k, _d(P_dk)))


def objective_fm(X, y, P_, w_, lams_, degree=2, loss="squared", regularizer="squaredl12", alpha=1,
                 beta=1, gamma=1, mean=False, fit_lower="explicit", fit_linear=True):
    """sum_i loss + alpha/2 |w|^2 + beta/2 |P|^2 + gamma Omega(P), the quantity the update rules of
    sparse_factorization_machines.py:175-353 minimise (alpha, beta, gamma times n when mean=True,
    :181-188).  Order o of P_ has degree `degree - o` (explicit lower orders, :207-225).
    Returns a dict of the parts and the total."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    X = _augment(X, fit_lower, fit_linear, degree)
    n = X.shape[0]
    a_, b_, g_ = (alpha * n, beta * n, gamma * n) if mean else (alpha, beta, gamma)
    y_pred = fm_output(X, P_, w_, lams_, degree, fit_linear, fit_lower)
    parts = {"loss": loss_sum(y_pred, y, loss),
             "l2_w": float(np.dot(w_, w_)) if fit_linear else 0.0,
             "l2_P": float(sum(np.sum(P_[o] ** 2) for o in range(P_.shape[0]))),
             "omega": float(sum(reg_eval(P_[o], regularizer, degree - o) for o in range(P_.shape[0])))}
    parts["total"] = parts["loss"] + 0.5 * a_ * parts["l2_w"] + 0.5 * b_ * parts["l2_P"] + g_ * parts["omega"]
    return parts


def objective_all_subsets(X, y, P_, lams_, loss="squared", regularizer="omegati", beta=1, gamma=1,
                          mean=False):
    """As objective_fm for sparse_all_subsets.py:80-201 (no linear term, degree = -1)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = X.shape[0]
    b_, g_ = (beta * n, gamma * n) if mean else (beta, gamma)
    y_pred = np.ascontiguousarray(poly_predict(X, P_, lams_, "all-subsets"))
    parts = {"loss": loss_sum(y_pred, y, loss), "l2_w": 0.0, "l2_P": float(np.sum(P_ ** 2)),
             "omega": reg_eval(P_, regularizer, -1)}
    parts["total"] = parts["loss"] + 0.5 * b_ * parts["l2_P"] + g_ * parts["omega"]
    return parts
