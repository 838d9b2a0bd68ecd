// Per-nonzero arithmetic shared by the cluster sweep (pcd.cu) and the window sweep (pcd_window.cu):
// gradient / curvature terms and the write-back of one sample record {y_pred, y, A^1..A^{m-1}}.
#pragma once
#include "common.cuh"

enum { KIND_LINEAR = 0, KIND_FM = 1, KIND_ALL = 2 };

template <int R> __device__ __forceinline__ void load_rec(const double *p, double (&r)[R]) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
    for (int u = 0; u < R / 2; u++) { double2 v = q[u]; r[2 * u] = v.x; r[2 * u + 1] = v.y; }
    if (R & 1) r[R - 1] = p[R - 1];
}

// per-nonzero gradient terms.  r = {y_pred, y, A^1.. }.  dA[] keeps the chain for the write-back.
template <int KIND, int DEG, int LOSS, int R, int ND>
__device__ __forceinline__ void nz_terms(const double (&r)[R], double x, double pold, double (&dA)[ND],
                                         double &tg, double &th) {
    const double dl = sp_dloss<LOSS>(r[0], r[1]);
    if (KIND == KIND_LINEAR) {
        dA[0] = x;
        tg += dl * x;                                        // cd_linear.py:18
    } else if (KIND == KIND_FM) {
        dA[0] = x;                                           // pcd.py:8-12
#pragma unroll
        for (int t = 1; t < DEG; t++) dA[t] = x * (r[1 + t] - pold * dA[t - 1]);
        tg += dl * dA[DEG - 1];                              // pcd.py:56-57
        th += dA[DEG - 1] * dA[DEG - 1];
    } else {
        dA[0] = x * r[2] / (1.0 + x * pold);                 // pcd_all.py:29-31
        tg += dl * dA[0];
        th += dA[0] * dA[0];
    }
}

// write-back of one sample after the coordinate moved by upd = p_old - p_new
template <int KIND, int DEG, int R, int ND>
__device__ __forceinline__ void nz_scatter(double *p, const double (&r)[R], const double (&dA)[ND], double x,
                                           double lam, double upd, double pold, double pnew) {
    if (KIND == KIND_LINEAR) {
        p[0] = r[0] - upd * x;                               // cd_linear.py:31
    } else if (KIND == KIND_FM) {
#pragma unroll
        for (int t = 1; t < DEG; t++) p[1 + t] = r[1 + t] - upd * dA[t - 1];   // pcd.py:129-130
        p[0] = r[0] - (lam * upd) * dA[DEG - 1];             // pcd.py:133
    } else {
        double yp = r[0] - lam * r[2];                       // pcd_all.py:95-98
        double A = r[2] / (1.0 + x * pold);
        A = A * (1.0 + x * pnew);
        yp = yp + lam * A;
        p[2] = A;
        p[0] = yp;
    }
}


// the chain dA[0..ND) alone (what the reference's synchronize loop recomputes, pcd.py:124-128)
template <int KIND, int DEG, int R, int ND>
__device__ __forceinline__ void nz_dA(const double (&r)[R], double x, double pold, double (&dA)[ND]) {
    dA[0] = x;
    if (KIND == KIND_FM) {
#pragma unroll
        for (int t = 1; t < DEG; t++) dA[t] = x * (r[1 + t] - pold * dA[t - 1]);
    }
}

// nz_scatter on a register copy of the record (r is updated in place; r[1] = y never changes)
template <int KIND, int DEG, int R, int ND>
__device__ __forceinline__ void nz_update(double (&r)[R], const double (&dA)[ND], double x, double lam,
                                          double upd, double pold, double pnew) {
    if (KIND == KIND_LINEAR) {
        r[0] = r[0] - upd * x;                               // cd_linear.py:31
    } else if (KIND == KIND_FM) {
#pragma unroll
        for (int t = 1; t < DEG; t++) r[1 + t] = r[1 + t] - upd * dA[t - 1];   // pcd.py:129-130
        r[0] = r[0] - (lam * upd) * dA[DEG - 1];             // pcd.py:133
    } else {
        double yp = r[0] - lam * r[2];                       // pcd_all.py:95-98
        double A = r[2] / (1.0 + x * pold);
        A = A * (1.0 + x * pnew);
        yp = yp + lam * A;
        r[2] = A;
        r[0] = yp;
    }
}

// prox_cd + incremental regularizer cache of one coordinate (l1.py:32-33, squaredl12.py:47-57,
// omegati.py:82-104).  p = p_old - eta*u (pre-prox), strength = eta*gamma/inv_step_size.
// cache: squaredl12 {||p_s||_1}; omegati FM {e_0..e_{m-1}} of |p_s|; omegati all-subsets {prod}.
template <int KIND, int DEG, int NC>
__device__ __forceinline__ double prox_chain(int reg, double p, double strength, double pold,
                                             double (&cache)[NC]) {
    double pnew;
    const double a_old = fabs(pold);
    if (reg == SP_REG_L1) {
        pnew = sp_soft_threshold(p, strength);
    } else if (reg == SP_REG_SQL12) {
        const double dcache = cache[0] - a_old;
        p = p / (1.0 + 2.0 * strength);
        const double sign = p > 0.0 ? 1.0 : -1.0;
        double m = fabs(p) - 2.0 * strength * dcache / (1.0 + 2.0 * strength);
        if (!(m > 0.0)) m = 0.0;
        pnew = sign * m;
        cache[0] = cache[0] - a_old;
        cache[0] = cache[0] + fabs(pnew);
    } else {
        const double sign = p > 0.0 ? 1.0 : -1.0;
        if (KIND == KIND_FM) {
            double dc[DEG + 1];
            dc[0] = 0.0; dc[1] = 1.0;
#pragma unroll
            for (int deg = 2; deg <= DEG; deg++) {
                double v = cache[deg - 1];
                v = v - dc[deg - 1] * a_old;
                if (v < 0.0) v = 0.0;
                dc[deg] = v;
            }
            strength = strength * dc[DEG];
            double m = fabs(p) - strength;
            if (!(m > 0.0)) m = 0.0;
            pnew = sign * m;
            const double a_new = fabs(pnew);
#pragma unroll
            for (int deg = 1; deg < DEG; deg++) cache[deg] = dc[deg + 1] + dc[deg] * a_new;
        } else {
            cache[0] = cache[0] / (1.0 + a_old);
            strength = strength * cache[0];
            double m = fabs(p) - strength;
            if (!(m > 0.0)) m = 0.0;
            pnew = sign * m;
            cache[0] = cache[0] * (1.0 + fabs(pnew));
        }
    }
    return pnew;
}
