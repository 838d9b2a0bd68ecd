// Regularizer cache kernels shared by pcd and pbcd.
#include "common.cuh"

namespace {
// ------------------------------------------------------------------ regularizer cache kernels
// compute_cache_pcd for one component (squaredl12.py:42-45, omegati.py:62-80):
//   mode 0 (squaredl12): regstate[0] = sum_j |p_j|
//   mode 1 (omegati FM): regstate[t] = e_t(|p_1|..|p_d|), t = 0..degree  (truncated product of
//                        the polynomials 1 + |p_j| z; the reference folds them left to right,
//                        here each thread folds a contiguous chunk and the chunks are multiplied
//                        in a fixed tree -- all terms are >= 0, so this is well conditioned)
//   mode 2 (omegati all-subsets): regstate[0] = prod_j (1 + |p_j|)
constexpr int RC_THREADS = 1024;
__global__ void __launch_bounds__(RC_THREADS) reg_cache_kernel(int mode, int degree, int d,
                                                               const double *__restrict__ p,
                                                               double *regstate) {
    __shared__ double sh[RC_THREADS][SP_MAXDEG + 1];
    const int tid = threadIdx.x;
    const int m = mode == 1 ? degree : 0;
    double e[SP_MAXDEG + 1];
    for (int t = 0; t <= SP_MAXDEG; t++) e[t] = 0.0;
    e[0] = (mode == 0) ? 0.0 : 1.0;
    const int chunk = (d + RC_THREADS - 1) / RC_THREADS;
    const int lo = tid * chunk, hi = min(d, lo + chunk);
    for (int j = lo; j < hi; j++) {
        const double av = fabs(p[j]);
        if (mode == 0) e[0] += av;
        else if (mode == 2) e[0] *= 1.0 + av;
        else
            for (int t = m; t >= 1; t--) e[t] += e[t - 1] * av;
    }
    for (int t = 0; t <= SP_MAXDEG; t++) sh[tid][t] = e[t];
    __syncthreads();
    for (int off = 1; off < RC_THREADS; off <<= 1) {        // fixed tree, left operand first
        if ((tid & (2 * off - 1)) == 0) {
            double l[SP_MAXDEG + 1], r[SP_MAXDEG + 1], o[SP_MAXDEG + 1];
            for (int t = 0; t <= SP_MAXDEG; t++) { l[t] = sh[tid][t]; r[t] = sh[tid + off][t]; }
            if (mode == 0) o[0] = l[0] + r[0];
            else if (mode == 2) o[0] = l[0] * r[0];
            else
                for (int t = 0; t <= m; t++) {
                    double acc = 0.0;
                    for (int u = 0; u <= t; u++) acc += l[u] * r[t - u];
                    o[t] = acc;
                }
            for (int t = 0; t <= m; t++) sh[tid][t] = o[t];
        }
        __syncthreads();
    }
    if (tid == 0)
        for (int t = 0; t <= SP_MAXDEG; t++) regstate[t] = (t <= m) ? sh[0][t] : 0.0;
}

}  // namespace

int sp_launch_reg_cache(int mode, int degree, int d, const double *v, double *regstate, cudaStream_t st) {
    if (mode == 1 && (degree < 1 || degree > SP_MAXDEG)) {
        sp_set_error("regularizer cache: degree %d unsupported", degree);
        return SP_ERR_UNSUPPORTED;
    }
    sp_prof_begin(SP_PROF_REGCACHE, st);
    reg_cache_kernel<<<1, RC_THREADS, 0, st>>>(mode, degree, d, v, regstate);
    sp_prof_end(st);
    SP_LAUNCH_CHECK("reg_cache_kernel");
    return SP_OK;
}
