// Helpers shared by the cluster block sweep (pbcd.cu) and the window block sweep (pbcd_window.cu).
#pragma once
#include "common.cuh"

enum { PK_FM = 1, PK_ALL = 2 };

// e_t(norms) for t=0..deg computed identically by every warp (deterministic: strided fold per
// lane, butterfly of truncated polynomial products, lane-0 broadcast).  `skip` is treated as 0.
template <int MAXD>
__device__ void warp_esp(const double *norms, int d, int skip, int deg, double (&out)[MAXD + 1]) {
    const int lane = threadIdx.x & 31;
    double e[MAXD + 1];
#pragma unroll
    for (int t = 0; t <= MAXD; t++) e[t] = (t == 0) ? 1.0 : 0.0;
    for (int j = lane; j < d; j += 32) {
        const double v = (j == skip) ? 0.0 : norms[j];
#pragma unroll
        for (int t = MAXD; t >= 1; t--)
            if (t <= deg) e[t] += e[t - 1] * v;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        double o[MAXD + 1], r[MAXD + 1];
#pragma unroll
        for (int t = 0; t <= MAXD; t++) o[t] = sp_shfl_xor(e[t], m);
#pragma unroll
        for (int t = 0; t <= MAXD; t++) {
            double acc = 0.0;
#pragma unroll
            for (int u = 0; u <= MAXD; u++)
                if (u <= t) acc += e[u] * o[t - u];
            r[t] = acc;
        }
#pragma unroll
        for (int t = 0; t <= MAXD; t++) e[t] = r[t];
    }
#pragma unroll
    for (int t = 0; t <= MAXD; t++) out[t] = (t <= deg) ? sp_shfl(e[t], 0) : 0.0;
}

static __device__ double warp_sum_array(const double *v, int d) {
    double acc = 0.0;
    for (int j = (threadIdx.x & 31); j < d; j += 32) acc += v[j];
    acc = sp_warp_allsum(acc);
    return sp_shfl(acc, 0);
}

