// Row-wise (CSR) dynamic-programming kernels: ANOVA degree-m / all-subsets kernel values per
// (sample, component).  Replaces, on device:
//   * kernels.poly_predict / anova_kernel / _all_subsets_fast   (reference kernels.py:71-153)
//   * pcd._precompute_A_all_degree / pcd_all._precompute_A_all  (optimizer/pcd.py:15-30,
//     optimizer/pcd_all.py:8-18)  -- one component, written into the per-sample record array
//   * pbcd._precompute_A_all_degree / pbcd_all._precompute_A_all (optimizer/pbcd.py:18-33,
//     optimizer/pbcd_all.py:9-20) -- all components, written into A[n, m-1, k]
// The reference walks columns (CSC) and scatters into A; per sample that visits its features in
// ascending order, which is exactly the CSR row order used here, so every A[i,t] sees the same
// sequence of fp64 operations (bit-identical DP), while the loads are coalesced row streams.
#include "common.cuh"
#include "sparsepoly_b200.h"

namespace {

constexpr int ROWS_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// all-components kernel: one group of G lanes per sample, lanes over components.
//   MODE 0: out[i*out_stride] (+)= sum_j x w_j + sum_s lam_s * K_s(i)      (predict)
//   MODE 1: Aout[i, t-1, s] = A^t_s(i), t=1..DEG-1  (ANOVA)  /  Aout[i, s] = A_s(i) (all-subsets)
//   MODE 2: Aout[i, s] = K_s(i) (the Gram matrix of kernels.anova_kernel / all_subsets_kernel)
// PX=true  : A += (A*p)*x   (pcd.py:30, pbcd.py:33)
// PX=false : A += (A*x)*p   (psgd.py:44)
template <int DEG, int G, int MODE, bool PX>
__global__ void __launch_bounds__(ROWS_THREADS)
rows_all_kernel(int n, int k, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                const double *__restrict__ data, const double *__restrict__ P_dk,
                const double *__restrict__ lams, const double *__restrict__ w, double *out,
                int out_stride, int accumulate, double *Aout) {
    constexpr int ND = DEG > 0 ? DEG : 1;
    const int lane = threadIdx.x & (G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu
                                     : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
    const int groups_per_block = ROWS_THREADS / G;
    const int group = blockIdx.x * groups_per_block + threadIdx.x / G;
    const int n_groups = gridDim.x * groups_per_block;
    for (int i = group; i < n; i += n_groups) {
        const int st = indptr[i], en = indptr[i + 1];
        double ypred = 0.0;
        if (MODE == 0 && w != nullptr) {
            // linear term <w, x_i>; tree order (reference: safe_sparse_dot)
            double acc = 0.0;
            for (int e = st + lane; e < en; e += G) acc += data[e] * w[indices[e]];
#pragma unroll
            for (int m = G / 2; m > 0; m >>= 1) acc += __shfl_xor_sync(gmask, acc, m, G);
            ypred = acc;
        }
        for (int s0 = 0; s0 < k; s0 += G) {
            const int s = s0 + lane;
            const bool act = s < k;
            double A[ND + 1];
            A[0] = 1.0;
#pragma unroll
            for (int t = 1; t <= ND; t++) A[t] = (DEG > 0) ? 0.0 : 1.0;
            for (int base = st; base < en; base += G) {
                const int e = base + lane;
                int jl = 0;
                double xl = 0.0;
                if (e < en) { jl = indices[e]; xl = data[e]; }
                const int cnt = min(G, en - base);
                for (int q0 = 0; q0 < cnt; q0 += 8) {
                    double pv[8], xv[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {      // 8 independent row gathers in flight
                        const int q = q0 + u;
                        const int j = __shfl_sync(gmask, jl, q & (G - 1), G);
                        xv[u] = __shfl_sync(gmask, xl, q & (G - 1), G);
                        pv[u] = (q < cnt && act) ? P_dk[(size_t)j * k + s] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        if (q0 + u < cnt) {
                            if (DEG > 0) {
#pragma unroll
                                for (int t = 0; t < ND; t++) {
                                    const double a = A[ND - t - 1];
                                    const double inc = PX ? (a * pv[u]) * xv[u] : (a * xv[u]) * pv[u];
                                    A[ND - t] += inc;
                                }
                            } else {
                                // kernels.py:129 (1 + x*p); pcd_all.py:18 / pbcd_all.py:20 (1.0 + p*x)
                                A[1] *= 1.0 + (PX ? pv[u] * xv[u] : xv[u] * pv[u]);
                            }
                        }
                    }
                }
            }
            if (MODE == 0) {
                double v = act ? lams[s] * A[ND] : 0.0;
#pragma unroll
                for (int m = G / 2; m > 0; m >>= 1) v += __shfl_xor_sync(gmask, v, m, G);
                ypred += v;
            } else if (MODE == 2) {
                if (act) Aout[(size_t)i * k + s] = A[ND];
            } else if (act) {
                if (DEG > 0) {
#pragma unroll
                    for (int t = 1; t < ND; t++)
                        Aout[((size_t)i * (ND - 1) + (t - 1)) * k + s] = A[t];
                } else {
                    Aout[(size_t)i * k + s] = A[1];
                }
            }
        }
        if (MODE == 0 && lane == 0) {
            double *o = out + (size_t)i * out_stride;
            *o = accumulate ? (*o + ypred) : ypred;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// single-component kernel for pcd: 8 lanes per sample; lanes own nonzeros (coalesced idx/val
// load + parallel gather of p_s[j]); the recurrence itself is run redundantly by all lanes.
// Writes rec[i*stride + 2 + (t-1)] = A^t(i), t=1..DEG-1 (ANOVA) or rec[i*stride+2] = A(i).
template <int DEG>
__global__ void __launch_bounds__(ROWS_THREADS)
rows_one_kernel(int n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                const double *__restrict__ data, const double *__restrict__ p_s, double *rec,
                int rec_stride) {
    constexpr int G = 8;
    constexpr int ND = DEG > 0 ? DEG : 1;
    const int lane = threadIdx.x & (G - 1);
    const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~(G - 1));
    const int groups_per_block = ROWS_THREADS / G;
    const int group = blockIdx.x * groups_per_block + threadIdx.x / G;
    const int n_groups = gridDim.x * groups_per_block;
    for (int i = group; i < n; i += n_groups) {
        const int st = indptr[i], en = indptr[i + 1];
        double A[ND + 1];
        A[0] = 1.0;
#pragma unroll
        for (int t = 1; t <= ND; t++) A[t] = (DEG > 0) ? 0.0 : 1.0;
        for (int base = st; base < en; base += G) {
            const int e = base + lane;
            double pl = 0.0, xl = 0.0;
            if (e < en) { xl = data[e]; pl = p_s[indices[e]]; }
            const int cnt = min(G, en - base);
#pragma unroll
            for (int q = 0; q < G; q++) {
                const double p = __shfl_sync(gmask, pl, q, G);
                const double x = __shfl_sync(gmask, xl, q, G);
                if (q < cnt) {
                    if (DEG > 0) {
#pragma unroll
                        for (int t = 0; t < ND; t++) A[ND - t] += (A[ND - t - 1] * p) * x;  // pcd.py:30
                    } else {
                        A[1] *= 1.0 + p * x;                                               // pcd_all.py:18
                    }
                }
            }
        }
        if (lane == 0) {
            double *r = rec + (size_t)i * rec_stride + 2;
            if (DEG > 0) {
#pragma unroll
                for (int t = 1; t < ND; t++) r[t - 1] = A[t];
            } else {
                r[0] = A[1];
            }
        }
    }
}

int grid_for(int n, int per_block) {
    long long blocks = ((long long)n + per_block - 1) / per_block;
    const long long cap = 148LL * 8 * 4;   // a few waves of 8 resident CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <int DEG, int MODE, bool PX>
int launch_all(int n, int k, const int32_t *indptr, const int32_t *indices, const double *data,
               const double *P_dk, const double *lams, const double *w, double *out, int out_stride,
               int accumulate, double *Aout, cudaStream_t st) {
    sp_prof_begin(SP_PROF_ROWS, st);
    if (k <= 8) {
        rows_all_kernel<DEG, 8, MODE, PX><<<grid_for(n, ROWS_THREADS / 8), ROWS_THREADS, 0, st>>>(
            n, k, indptr, indices, data, P_dk, lams, w, out, out_stride, accumulate, Aout);
    } else if (k <= 16) {
        rows_all_kernel<DEG, 16, MODE, PX><<<grid_for(n, ROWS_THREADS / 16), ROWS_THREADS, 0, st>>>(
            n, k, indptr, indices, data, P_dk, lams, w, out, out_stride, accumulate, Aout);
    } else {
        rows_all_kernel<DEG, 32, MODE, PX><<<grid_for(n, ROWS_THREADS / 32), ROWS_THREADS, 0, st>>>(
            n, k, indptr, indices, data, P_dk, lams, w, out, out_stride, accumulate, Aout);
    }
    sp_prof_end(st);
    SP_LAUNCH_CHECK("rows_all_kernel");
    return SP_OK;
}

template <int MODE, bool PX>
int dispatch_all(int degree, int n, int k, const int32_t *indptr, const int32_t *indices,
                 const double *data, const double *P_dk, const double *lams, const double *w,
                 double *out, int out_stride, int accumulate, double *Aout, cudaStream_t st) {
    switch (degree) {
    case -1: return launch_all<-1, MODE, PX>(n, k, indptr, indices, data, P_dk, lams, w, out, out_stride, accumulate, Aout, st);
    case 1: return launch_all<1, MODE, PX>(n, k, indptr, indices, data, P_dk, lams, w, out, out_stride, accumulate, Aout, st);
    case 2: return launch_all<2, MODE, PX>(n, k, indptr, indices, data, P_dk, lams, w, out, out_stride, accumulate, Aout, st);
    case 3: return launch_all<3, MODE, PX>(n, k, indptr, indices, data, P_dk, lams, w, out, out_stride, accumulate, Aout, st);
    case 4: return launch_all<4, MODE, PX>(n, k, indptr, indices, data, P_dk, lams, w, out, out_stride, accumulate, Aout, st);
    case 5: return launch_all<5, MODE, PX>(n, k, indptr, indices, data, P_dk, lams, w, out, out_stride, accumulate, Aout, st);
    default:
        sp_set_error("degree %d is not supported (1..%d or -1 for all-subsets)", degree, SP_MAXDEG);
        return SP_ERR_UNSUPPORTED;
    }
}

}  // namespace

extern "C" int sp_predict(const sp_dataset *ds, const double *P_dk, int k, const double *lams,
                          int degree, const double *w, double *out, int out_stride, int accumulate,
                          sp_stream stream) {
    if (!ds || !ds->csr_indptr || !P_dk || !lams || !out || k <= 0 || out_stride <= 0) {
        sp_set_error("sp_predict: invalid argument");
        return SP_ERR_INVALID;
    }
    if (ds->n_samples == 0) return SP_OK;
    return dispatch_all<0, false>(degree, ds->n_samples, k, ds->csr_indptr, ds->csr_indices,
                                  ds->csr_data, P_dk, lams, w, out, out_stride, accumulate, nullptr,
                                  (cudaStream_t)stream);
}

extern "C" int sp_kernel_matrix(const sp_dataset *ds, const double *P_dk, int k, int degree, double *K,
                                sp_stream stream) {
    if (!ds || !ds->csr_indptr || !P_dk || !K || k <= 0) {
        sp_set_error("sp_kernel_matrix: invalid argument");
        return SP_ERR_INVALID;
    }
    if (ds->n_samples == 0) return SP_OK;
    return dispatch_all<2, false>(degree, ds->n_samples, k, ds->csr_indptr, ds->csr_indices,
                                  ds->csr_data, P_dk, nullptr, nullptr, nullptr, 1, 0, K,
                                  (cudaStream_t)stream);
}

// pbcd cache precompute (all components): A[n, m-1, k] (ANOVA) or A[n, k] (all-subsets)
int sp_rows_precompute_all(const sp_dataset *ds, const double *P_dk, int k, int degree, double *A,
                           cudaStream_t st) {
    if (ds->n_samples == 0) return SP_OK;
    return dispatch_all<1, true>(degree, ds->n_samples, k, ds->csr_indptr, ds->csr_indices,
                                 ds->csr_data, P_dk, nullptr, nullptr, nullptr, 1, 0, A, st);
}

// pcd cache precompute (one component) into the record array
int sp_rows_precompute_one(const sp_dataset *ds, const double *p_s, int degree, double *rec,
                           int rec_stride, cudaStream_t st) {
    const int n = ds->n_samples;
    if (n == 0) return SP_OK;
    const int grid = grid_for(n, ROWS_THREADS / 8);
#define SP_ONE(D)                                                                                \
    rows_one_kernel<D><<<grid, ROWS_THREADS, 0, st>>>(n, ds->csr_indptr, ds->csr_indices,        \
                                                      ds->csr_data, p_s, rec, rec_stride)
    sp_prof_begin(SP_PROF_ROWS, st);
    switch (degree) {
    case -1: SP_ONE(-1); break;
    case 2: SP_ONE(2); break;
    case 3: SP_ONE(3); break;
    case 4: SP_ONE(4); break;
    case 5: SP_ONE(5); break;
    default:
        sp_set_error("pcd degree %d is not supported (2..%d or -1)", degree, SP_MAXDEG);
        return SP_ERR_UNSUPPORTED;
    }
#undef SP_ONE
    sp_prof_end(st);
    SP_LAUNCH_CHECK("rows_one_kernel");
    return SP_OK;
}
