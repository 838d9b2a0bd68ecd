// Proximal minibatch SGD on device (reference optimizer/psgd.py:9-199) and the whole-matrix
// proximal operators it calls (regularizer/l1.py:50-51, l21.py:43-48, squaredl12.py:66-78,
// squaredl21.py:63-74, utils.py:26-70).
//
//   sp_psgd_grad : one group of lanes per sample, lanes over components; the sample's ANOVA DP
//                  (psgd._anova) runs in registers, P rows are gathered as coalesced k-vectors,
//                  gradients are scattered with fp64 RED atomics into the dense grad_P / grad_w.
//   sp_psgd_step : fused dense update  P = (P - c*G)/den ; G = 0   (psgd._update_params)
//   sp_prox      : l1 / l21 elementwise & row kernels; squared-l1,2 by a cooperative
//                  fixed-point selection kernel (Michelot-style): the active set
//                  G <- {i in G : |p_i| > 2*s*S_G/(1+2*s*|G|)} shrinks monotonically to the same
//                  (theta, S) the reference's randomized-pivot search finds.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "sparsepoly_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int PG_THREADS = 256;
constexpr int PG_HOT = SP_MAX_HOT_FEATURES;   // dense features whose gradient is pre-reduced per warp

// ------------------------------------------------------------------------------ gradient
// One group of G lanes per sample; lane l owns components l, l+G, ... (KCH of them).
template <int DEG, int NORD, int G, int KCH>
__global__ void __launch_bounds__(PG_THREADS)
psgd_grad_kernel(int k, int d, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                 const double *__restrict__ data, const double *__restrict__ y,
                 const double *__restrict__ P, const double *__restrict__ w,
                 const double *__restrict__ lams, int loss, int fit_linear,
                 const int32_t *__restrict__ idx_samples, int b0, int b1, double *grad_P,
                 double *grad_w, double *loss_sum,
                 const int8_t *__restrict__ feat_hot, int n_hot, const int32_t *__restrict__ hot_feat) {
    const int lane = threadIdx.x & (G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu
                                     : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
    const int groups_per_block = PG_THREADS / G;
    const int group = blockIdx.x * groups_per_block + threadIdx.x / G;
    const int n_groups = gridDim.x * groups_per_block;
    const size_t dk = (size_t)d * k;
    // Dense ("hot") features are hit by (almost) every sample: their same-address fp64 REDs
    // serialise in L2.  Each group of lanes therefore sums its samples' contributions to those
    // rows in a private shared-memory tile (lane = component, no atomics) and issues one RED per
    // (row, component) when it is done.
    constexpr bool HOT = (KCH * NORD == 1);
    extern __shared__ double pg_smem[];
    double *hacc = pg_smem + (size_t)(threadIdx.x / G) * (PG_HOT * G + PG_HOT);   // [PG_HOT][G] + w[PG_HOT]
    double *hw = hacc + PG_HOT * G;
    const bool use_hot = HOT && feat_hot != nullptr && n_hot > 0;
    if (use_hot) {
        for (int q = lane; q < PG_HOT * G + PG_HOT; q += G) hacc[q] = 0.0;
        __syncwarp(gmask);
    }
    double lam[KCH];
#pragma unroll
    for (int c = 0; c < KCH; c++) lam[c] = (lane + G * c < k) ? lams[lane + G * c] : 0.0;
    double loss_acc = 0.0;
    for (int b = b0 + group; b < b1; b += n_groups) {
        const int i = idx_samples[b];
        const int st = indptr[i], en = indptr[i + 1];
        // ---- _pred, psgd.py:47-57
        double ypred = 0.0;
        for (int e = st + lane; e < en; e += G) ypred += data[e] * w[indices[e]];
#pragma unroll
        for (int m = G / 2; m > 0; m >>= 1) ypred += __shfl_xor_sync(gmask, ypred, m, G);
        double A[KCH][NORD][DEG + 1];
#pragma unroll
        for (int c = 0; c < KCH; c++)
#pragma unroll
            for (int o = 0; o < NORD; o++) {
                A[c][o][0] = 1.0;
#pragma unroll
                for (int t = 1; t <= DEG; t++) A[c][o][t] = 0.0;
            }
        for (int base = st; base < en; base += G) {
            const int e = base + lane;
            int jl = 0;
            double xl = 0.0;
            if (e < en) { jl = indices[e]; xl = data[e]; }
            const int cnt = min(G, en - base);
            constexpr int UB = (KCH * NORD <= 2) ? 4 : 2;
            for (int q0 = 0; q0 < cnt; q0 += UB) {
                double pv[UB][KCH][NORD], xv[UB];
#pragma unroll
                for (int u = 0; u < UB; u++) {            // UB independent row gathers in flight
                    const int q = q0 + u;
                    const int j = __shfl_sync(gmask, jl, q & (G - 1), G);
                    xv[u] = __shfl_sync(gmask, xl, q & (G - 1), G);
#pragma unroll
                    for (int c = 0; c < KCH; c++)
#pragma unroll
                        for (int o = 0; o < NORD; o++)
                            pv[u][c][o] = (q < cnt && lane + G * c < k) ? P[o * dk + (size_t)j * k + lane + G * c] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    if (q0 + u < cnt) {
#pragma unroll
                        for (int c = 0; c < KCH; c++)
#pragma unroll
                            for (int o = 0; o < NORD; o++)
#pragma unroll
                                for (int t = 0; t < DEG - o; t++)       // _anova, psgd.py:34-44
                                    A[c][o][DEG - o - t] += (A[c][o][DEG - o - t - 1] * xv[u]) * pv[u][c][o];
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < NORD; o++) {                  // y_pred += dot(lams, A[order, deg])
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < KCH; c++) v += lam[c] * A[c][o][DEG - o];
#pragma unroll
            for (int m = G / 2; m > 0; m >>= 1) v += __shfl_xor_sync(gmask, v, m, G);
            ypred += v;
        }
        const double yi = y[i];
        if (lane == 0) loss_acc += sp_loss_rt(loss, ypred, yi);     // psgd.py:155
        const double dL = sp_dloss_rt(loss, ypred, yi);
        // ---- _update_grads, psgd.py:60-91
        if (fit_linear)
            for (int e = st + lane; e < en; e += G) {
                const int j = indices[e];
                const int h = use_hot ? (int)feat_hot[j] : -1;
                if (h >= 0) hw[h] += dL * data[e];               // (a row's features are distinct)
                else atomicAdd(grad_w + j, dL * data[e]);
            }
        if (use_hot) __syncwarp(gmask);                          // hw[] is shared by the group's lanes
        for (int base = st; base < en; base += G) {
            const int e = base + lane;
            int jl = 0, hl = -1;
            double xl = 0.0;
            if (e < en) { jl = indices[e]; xl = data[e]; if (use_hot) hl = (int)feat_hot[jl]; }
            const int cnt = min(G, en - base);
#pragma unroll 4
            for (int q = 0; q < cnt; q++) {
                const int j = __shfl_sync(gmask, jl, q, G);
                const double x = __shfl_sync(gmask, xl, q, G);
                const int hq = HOT ? __shfl_sync(gmask, hl, q, G) : -1;
#pragma unroll
                for (int c = 0; c < KCH; c++) {
                    const int s = lane + G * c;
                    if (s < k) {
#pragma unroll
                        for (int o = 0; o < NORD; o++) {
                            double dprev = x;                           // _grad_anova, psgd.py:25-31
                            if (DEG - o > 1) {
                                const double p = P[o * dk + (size_t)j * k + s];
#pragma unroll
                                for (int t = 1; t < DEG - o; t++) dprev = x * (A[c][o][t] - p * dprev);
                            }
                            const double gv = (dL * lam[c]) * dprev;                    // psgd.py:91
                            if (HOT && hq >= 0) hacc[hq * G + lane] += gv;
                            else atomicAdd(grad_P + o * dk + (size_t)j * k + s, gv);
                        }
                    }
                }
            }
        }
    }
    if (use_hot) {
        __syncwarp(gmask);
        for (int h = 0; h < n_hot; h++) {
            const int j = hot_feat[h];
            const double gv = hacc[h * G + lane];
            if (lane < k && gv != 0.0) atomicAdd(grad_P + (size_t)j * k + lane, gv);
            if (lane == 0 && hw[h] != 0.0) atomicAdd(grad_w + j, hw[h]);
        }
    }
    if (lane == 0 && loss_acc != 0.0) atomicAdd(loss_sum, loss_acc);
}

// ------------------------------------------------------------------------------ dense step
__global__ void psgd_step_kernel(double *__restrict__ P, double *__restrict__ G, size_t n, double c,
                                 double den) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n2 = n / 2;
    double2 *P2 = reinterpret_cast<double2 *>(P);
    double2 *G2 = reinterpret_cast<double2 *>(G);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += stride) {
        double2 p = P2[e], g = G2[e];
        g.x = g.x * c; g.y = g.y * c;                 // grad *= eta / batch      psgd.py:113
        p.x = p.x - g.x; p.y = p.y - g.y;             // P -= grad                psgd.py:114
        p.x = p.x / den; p.y = p.y / den;             // P /= 1 + eta*beta        psgd.py:115
        P2[e] = p;
        G2[e] = make_double2(0.0, 0.0);               // grad[:] = 0              psgd.py:195
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) {
        double g = G[n - 1] * c;
        P[n - 1] = (P[n - 1] - g) / den;
        G[n - 1] = 0.0;
    }
}

// ------------------------------------------------------------------------------ prox: l1, l21
__global__ void prox_l1_kernel(double *__restrict__ P, size_t n, double strength) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride)
        P[e] = sp_soft_threshold(P[e], strength);
}

// l21.py:43-48.  Reference quirk kept: rows with norm <= strength get factor 1 - s/inf = 1,
// i.e. they are left unchanged (not zeroed).
__global__ void prox_l21_kernel(double *__restrict__ P, int d, int k, double strength) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < d; j += n_warps) {
        double acc = 0.0;
        for (int s = lane; s < k; s += 32) { const double v = fabs(P[(size_t)j * k + s]); acc += v * v; }
        acc = sp_warp_allsum(acc);
        const double nm = pow(acc, 0.5);
        if (nm > strength) {
            const double sc = 1.0 - strength / nm;
            for (int s = lane; s < k; s += 32) P[(size_t)j * k + s] *= sc;
        }
    }
}

// ------------------------------------------------------------------------------ squared-l1,2
// Cooperative fixed-point selection over the columns of V[rows, cols] (|.| taken on the fly).
//   state per column: tau (current threshold), theta (active count)
//   repeat: (cnt, sum) over {|v| > tau}  ->  tau' = 2*s*sum / (1 + 2*s*cnt)
//   until no column's count changes.  Then S = sum/(1+2*s*cnt) and the caller soft-thresholds
//   with 2*s*S (utils.py:69-70).
// work layout (doubles): tau[cols] | S[cols] | partial_sum[nblk*cols] | partial_cnt[nblk*cols]
struct SelArgs {
    const double *V;
    int rows, cols;
    double strength;
    double *tau, *S, *psum, *pcnt;
    int *flags;      // [0] changed-counter
    int max_iter;
};

__global__ void __launch_bounds__(256) sql12_select_kernel(const SelArgs a) {
    cg::grid_group grid = cg::this_grid();
    const int cols = a.cols, rows = a.rows;
    const int nblk = gridDim.x, tid = threadIdx.x, T = blockDim.x;
    extern __shared__ double sh[];          // [T] sums, [T] counts
    double *ssum = sh, *scnt = sh + T;
    // thread <-> column mapping: consecutive threads walk consecutive columns of a row
    // (coalesced); rows_per_pass rows are covered by one block pass.
    const int tpr = cols < T ? cols : T;             // threads per row (cols <= T assumed, else loop)
    const int rpp = T / tpr;                          // rows per block pass
    const int my_col0 = tid % tpr, my_row0 = tid / tpr;
    const bool worker = tid < tpr * rpp;
    for (int c = blockIdx.x * T + tid; c < cols; c += nblk * T) a.tau[c] = -1.0;   // all |v|>=0 active, zeros dropped below
    if (blockIdx.x == 0 && tid == 0) a.flags[0] = 0;
    grid.sync();
    double prev_cnt_total = -1.0;
    for (int it = 0; it < a.max_iter; it++) {
        // ---- phase 1: per-block partial (sum, cnt) for each column
        for (int c0 = 0; c0 < cols; c0 += tpr) {
            const int col = c0 + my_col0;
            double lsum = 0.0, lcnt = 0.0;
            if (worker && col < cols) {
                const double tau = a.tau[col];
                for (long long r = (long long)blockIdx.x * rpp + my_row0; r < rows; r += (long long)nblk * rpp) {
                    const double v = fabs(a.V[(size_t)r * cols + col]);
                    if (v > tau && v > 0.0) { lsum += v; lcnt += 1.0; }
                }
            }
            ssum[tid] = lsum; scnt[tid] = lcnt;
            __syncthreads();
            if (tid < tpr && c0 + tid < cols) {          // fixed-order combine over the block's rows
                double s = 0.0, n = 0.0;
                for (int r = 0; r < rpp; r++) { s += ssum[r * tpr + tid]; n += scnt[r * tpr + tid]; }
                a.psum[(size_t)blockIdx.x * cols + c0 + tid] = s;
                a.pcnt[(size_t)blockIdx.x * cols + c0 + tid] = n;
            }
            __syncthreads();
        }
        grid.sync();
        // ---- phase 2: every block redundantly reduces the partials of its columns (fixed order)
        double cnt_total = 0.0;
        for (int col = tid; col < cols; col += T) {
            double s = 0.0, n = 0.0;
            for (int b = 0; b < nblk; b++) { s += a.psum[(size_t)b * cols + col]; n += a.pcnt[(size_t)b * cols + col]; }
            const double den = 1.0 + 2.0 * a.strength * n;
            if (blockIdx.x == 0) {
                a.tau[col] = 2.0 * a.strength * s / den;
                a.S[col] = s / den;
            }
            cnt_total += n;
        }
        // block-wide total count (identical in every block) decides convergence
        ssum[tid] = cnt_total;
        __syncthreads();
        for (int off = T / 2; off > 0; off >>= 1) { if (tid < off) ssum[tid] += ssum[tid + off]; __syncthreads(); }
        const double total = ssum[0];
        __syncthreads();
        grid.sync();                                   // tau visible to all blocks
        if (total == prev_cnt_total) break;            // no column lost an element: fixed point
        prev_cnt_total = total;
    }
}

__global__ void soft_threshold_cols_kernel(double *__restrict__ P, size_t n, int cols, double strength,
                                           const double *__restrict__ S) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const double th = 2.0 * strength * S[e % cols];
        P[e] = sp_soft_threshold(P[e], th);
    }
}

// squaredl21.py:63-74 helpers
__global__ void row_norm_pow_kernel(const double *__restrict__ P, int d, int k, double *norms) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < d; j += n_warps) {
        double acc = 0.0;
        for (int s = lane; s < k; s += 32) { const double v = fabs(P[(size_t)j * k + s]); acc += v * v; }
        acc = sp_warp_allsum(acc);
        if (lane == 0) norms[j] = pow(acc, 0.5);
    }
}
__global__ void row_rescale_kernel(double *__restrict__ P, int d, int k, const double *__restrict__ norms,
                                   double strength, const double *__restrict__ S) {
    const size_t n = (size_t)d * k, stride = (size_t)gridDim.x * blockDim.x;
    const double th = 2.0 * strength * S[0];
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const double nm = norms[e / k];
        double v = P[e];
        if (nm > 0.0) v = v / nm;                      // P[idx] /= norms[idx]
        v = v * sp_soft_threshold(nm, th);             // P *= prox(norms)
        P[e] = v;
    }
}


int ew_blocks(size_t n) {
    size_t b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (int)b;
}

int run_select(const double *V, int rows, int cols, double strength, double *work, cudaStream_t st,
               double **S_out) {
    if (cols > 256) { sp_set_error("squared-l1,2 prox: more than 256 columns (n_components) is not supported"); return SP_ERR_UNSUPPORTED; }
    int dev = 0, sms = 0, occ = 0;
    SP_CUDA(cudaGetDevice(&dev));
    SP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t shmem = 2 * 256 * sizeof(double);
    SP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sql12_select_kernel, 256, shmem));
    if (occ < 1) { sp_set_error("squared-l1,2 prox: kernel does not fit"); return SP_ERR_CUDA; }
    if (occ > 4) occ = 4;
    int nblk = sms * occ;
    const int tpr = cols < 256 ? cols : 256, rpp = 256 / tpr;
    const long long need = ((long long)rows + rpp - 1) / rpp;
    if (nblk > need) nblk = (int)(need < 1 ? 1 : need);
    SelArgs a;
    a.V = V; a.rows = rows; a.cols = cols; a.strength = strength;
    a.tau = work; a.S = work + cols;
    a.psum = work + 2 * (size_t)cols;
    a.pcnt = a.psum + (size_t)nblk * cols;
    a.flags = reinterpret_cast<int *>(a.pcnt + (size_t)nblk * cols);
    a.max_iter = 200;
    void *args[] = {(void *)&a};
    SP_CUDA(cudaLaunchCooperativeKernel((void *)sql12_select_kernel, dim3(nblk), dim3(256), args, shmem, st));
    *S_out = a.S;
    return SP_OK;
}

}  // namespace

extern "C" size_t sp_prox_work_doubles(int d, int k) {
    // norms[d] (squaredl21) + tau/S [2*max(k,1)] + partials 2*nblk*cols (nblk <= 148*4) + flags
    const size_t cols = (size_t)(k > 1 ? k : 1);
    return (size_t)d + 2 * cols + 2 * (size_t)148 * 4 * cols + 64;
}

extern "C" int sp_get_eta(int lr, double eta0, double alpha, double beta, double power_t, int64_t it,
                          double *eta_P, double *eta_w) {
    if (!eta_P || !eta_w) { sp_set_error("sp_get_eta: null pointer"); return SP_ERR_INVALID; }
    switch (lr) {                                       // psgd.py:9-22
    case 0: *eta_P = eta0; *eta_w = eta0; break;
    case 1: {
        const double eta_it = eta0 * (double)it;
        *eta_P = eta0 / pow(1.0 + eta_it * beta, power_t);
        *eta_w = eta0 / pow(1.0 + eta_it * alpha, power_t);
        break;
    }
    case 2: *eta_P = 1.0 / (beta * (double)it); *eta_w = 1.0 / (alpha * (double)it); break;
    case 3: { const double e = eta0 / pow((double)it, power_t); *eta_P = e; *eta_w = e; break; }
    default: sp_set_error("learning_rate id %d is not supported", lr); return SP_ERR_INVALID;
    }
    return SP_OK;
}

template <int DEG, int NORD>
static int launch_grad(cudaStream_t st, int k, int d, const sp_dataset *ds, const double *y,
                       const double *P, const double *w, const double *lams, int loss, int fit_linear,
                       const int32_t *idx, int b0, int b1, double *gP, double *gw, double *ls) {
    const int G = k <= 16 ? 16 : 32;
    const int per_block = PG_THREADS / G;
    long long blocks = ((long long)(b1 - b0) + per_block - 1) / per_block;
    if (blocks > 148LL * 8 * 8) blocks = 148LL * 8 * 8;
    const size_t smem = (size_t)(PG_THREADS / G) * (PG_HOT * G + PG_HOT) * sizeof(double);
    // persistent grid: the fewer groups, the fewer flushes of the hot-feature tiles
    if (ds->feat_hot && ds->n_hot_feat > 0 && blocks > 148LL * 4) blocks = 148LL * 4;
#define SP_GRAD(GG, KC)                                                                           \
    {                                                                                             \
        const size_t sm = (KC * NORD == 1) ? smem : 0;                                            \
        if (sm) {                                                                                 \
            cudaError_t e_ = cudaFuncSetAttribute(psgd_grad_kernel<DEG, NORD, GG, KC>,            \
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
            if (e_ != cudaSuccess) return sp_check_cuda(e_, "cudaFuncSetAttribute(psgd_grad_kernel)"); \
        }                                                                                         \
        psgd_grad_kernel<DEG, NORD, GG, KC><<<(int)blocks, PG_THREADS, sm, st>>>(k, d, ds->csr_indptr, \
            ds->csr_indices, ds->csr_data, y, P, w, lams, loss, fit_linear, idx, b0, b1, gP, gw, ls, \
            ds->feat_hot, ds->n_hot_feat, ds->hot_feat); \
    }
    sp_prof_begin(SP_PROF_PSGD_GRAD, st);
    if (k <= 16) SP_GRAD(16, 1)
    else if (k <= 32) SP_GRAD(32, 1)
    else if (k <= 64) SP_GRAD(32, 2)
    else if (k <= 128) SP_GRAD(32, 4)
    else { sp_set_error("psgd: n_components=%d > 128 is not supported by the CUDA backend", k); return SP_ERR_UNSUPPORTED; }
#undef SP_GRAD
    sp_prof_end(st);
    SP_LAUNCH_CHECK("psgd_grad_kernel");
    return SP_OK;
}

extern "C" int sp_psgd_grad(const sp_dataset *ds, const double *y, const double *P_odk, int n_orders,
                            int k, const double *w, const double *lams, int degree, int loss,
                            int fit_linear, const int32_t *idx_samples, int b0, int b1,
                            double *grad_P, double *grad_w, double *loss_sum, sp_stream stream) {
    if (!ds || !ds->csr_indptr || !y || !P_odk || !w || !lams || !idx_samples || !grad_P || !grad_w ||
        !loss_sum || k <= 0 || b0 < 0 || b1 < b0) {
        sp_set_error("sp_psgd_grad: invalid argument");
        return SP_ERR_INVALID;
    }
    if (degree < 2 || degree > SP_MAXDEG) {
        sp_set_error("psgd degree %d is not supported by the CUDA backend (2..%d)", degree, SP_MAXDEG);
        return SP_ERR_UNSUPPORTED;
    }
    if (n_orders != 1 && n_orders != degree - 1) {
        sp_set_error("psgd: n_orders must be 1 or degree-1 (got %d)", n_orders);
        return SP_ERR_INVALID;
    }
    if (loss < 0 || loss > 2) { sp_set_error("unknown loss id %d", loss); return SP_ERR_INVALID; }
    if (b1 == b0) return SP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int d = ds->n_features;
    const bool ex = n_orders > 1;
#define SP_CALL(D, N) return launch_grad<D, N>(st, k, d, ds, y, P_odk, w, lams, loss, fit_linear, idx_samples, b0, b1, grad_P, grad_w, loss_sum)
    switch (degree) {
    case 2: SP_CALL(2, 1);
    case 3: if (ex) SP_CALL(3, 2); else SP_CALL(3, 1);
    case 4: if (ex) SP_CALL(4, 3); else SP_CALL(4, 1);
    case 5: if (ex) SP_CALL(5, 4); else SP_CALL(5, 1);
    }
#undef SP_CALL
    return SP_ERR_UNSUPPORTED;
}

extern "C" int sp_psgd_step(double *P_odk, double *grad_P, double *w, double *grad_w, int n_orders,
                            int d, int k, double eta_P, double eta_w, double alpha, double beta,
                            int batch, int fit_linear, sp_stream stream) {
    if (((!P_odk || !grad_P) && n_orders > 0) || !w || !grad_w || batch <= 0) {
        sp_set_error("sp_psgd_step: invalid argument");
        return SP_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    sp_prof_begin(SP_PROF_PSGD_STEP, st);
    if (fit_linear && d > 0) {                                      // psgd.py:109-112
        psgd_step_kernel<<<ew_blocks((size_t)d / 2 + 1), 256, 0, st>>>(w, grad_w, (size_t)d, eta_w / batch,
                                                                        1 + eta_w * alpha);
        SP_LAUNCH_CHECK("psgd_step_kernel(w)");
    }
    const size_t n = (size_t)n_orders * d * k;
    if (n > 0) {
        psgd_step_kernel<<<ew_blocks(n / 2 + 1), 256, 0, st>>>(P_odk, grad_P, n, eta_P / batch,
                                                                1.0 + eta_P * beta);
        SP_LAUNCH_CHECK("psgd_step_kernel(P)");
    }
    sp_prof_end(st);
    return SP_OK;
}

extern "C" int sp_prox(double *P_dk, int d, int k, int reg, double strength, double *work,
                       sp_stream stream) {
    if (!P_dk || d < 0 || k <= 0) { sp_set_error("sp_prox: invalid argument"); return SP_ERR_INVALID; }
    if (d == 0) return SP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)d * k;
    struct ProfScope { cudaStream_t s; ProfScope(cudaStream_t s_) : s(s_) { sp_prof_begin(SP_PROF_PROX, s); } ~ProfScope() { sp_prof_end(s); } } prof_scope(st);
    switch (reg) {
    case SP_REG_L1:
        prox_l1_kernel<<<ew_blocks(n), 256, 0, st>>>(P_dk, n, strength);
        SP_LAUNCH_CHECK("prox_l1_kernel");
        return SP_OK;
    case SP_REG_L21: {
        int blocks = (d + 7) / 8; if (blocks > 148 * 16) blocks = 148 * 16;
        prox_l21_kernel<<<blocks, 256, 0, st>>>(P_dk, d, k, strength);
        SP_LAUNCH_CHECK("prox_l21_kernel");
        return SP_OK;
    }
    case SP_REG_SQL12: {
        if (!work) { sp_set_error("sp_prox: squaredl12 needs a work buffer"); return SP_ERR_INVALID; }
        double *S = nullptr;
        int rc = run_select(P_dk, d, k, strength, work, st, &S);
        if (rc) return rc;
        soft_threshold_cols_kernel<<<ew_blocks(n), 256, 0, st>>>(P_dk, n, k, strength, S);
        SP_LAUNCH_CHECK("soft_threshold_cols_kernel");
        return SP_OK;
    }
    case SP_REG_SQL21: {
        if (!work) { sp_set_error("sp_prox: squaredl21 needs a work buffer"); return SP_ERR_INVALID; }
        double *norms = work;
        int blocks = (d + 7) / 8; if (blocks > 148 * 16) blocks = 148 * 16;
        row_norm_pow_kernel<<<blocks, 256, 0, st>>>(P_dk, d, k, norms);
        SP_LAUNCH_CHECK("row_norm_pow_kernel");
        double *S = nullptr;
        int rc = run_select(norms, d, 1, strength, work + d, st, &S);
        if (rc) return rc;
        row_rescale_kernel<<<ew_blocks(n), 256, 0, st>>>(P_dk, d, k, norms, strength, S);
        SP_LAUNCH_CHECK("row_rescale_kernel");
        return SP_OK;
    }
    default:
        sp_set_error("regularizer id %d does not implement the psgd prox (use l1, l21, squaredl12 or squaredl21)", reg);
        return SP_ERR_UNSUPPORTED;
    }
}


extern "C" int sp_psgd_epoch(const sp_dataset *ds, const double *y, double *P_odk, int n_orders, int k,
                             double *w, const double *lams, int degree, double alpha, double beta,
                             double gamma, int reg, int loss, double *grad_P, double *grad_w,
                             const int32_t *idx_samples, int fit_linear, double eta0,
                             int learning_rate, double power_t, int batch_size, int64_t *it_io_host,
                             double *loss_sum, double *work, sp_stream stream) {
    if (!ds || !it_io_host || batch_size <= 0 || !work) { sp_set_error("sp_psgd_epoch: invalid argument"); return SP_ERR_INVALID; }
    const int n = ds->n_samples, d = ds->n_features;
    int64_t it = *it_io_host;
    for (int b0 = 0; b0 < n; b0 += batch_size) {           // psgd.py:150-198
        const int b1 = (n - b0 < batch_size) ? n : b0 + batch_size;
        int rc = sp_psgd_grad(ds, y, P_odk, n_orders, k, w, lams, degree, loss, fit_linear, idx_samples,
                              b0, b1, grad_P, grad_w, loss_sum, stream);
        if (rc) return rc;
        double eta_P, eta_w;
        rc = sp_get_eta(learning_rate, eta0, alpha, beta, power_t, it, &eta_P, &eta_w);
        if (rc) return rc;
        const double strength = gamma * eta_P / (1 + eta_P * beta);   // psgd.py:122
        rc = sp_psgd_step(P_odk, grad_P, w, grad_w, n_orders, d, k, eta_P, eta_w, alpha, beta, b1 - b0,
                          fit_linear, stream);
        if (rc) return rc;
        for (int o = 0; o < n_orders; o++) {                // psgd.py:119-122
            rc = sp_prox(P_odk + (size_t)o * d * k, d, k, reg, strength, work, stream);
            if (rc) return rc;
        }
        it++;
    }
    *it_io_host = it;
    return SP_OK;
}
