// Proximal BLOCK coordinate descent sweeps on one thread-block cluster (sm_100a):
//   * pbcd.pbcd_epoch      (reference optimizer/pbcd.py:36-148)
//   * pbcd_all.pbcd_epoch  (reference optimizer/pbcd_all.py:23-132)
// One block step = all k components of feature j.  Lanes run over components (coalesced rows of
// A[i, t, :] and P[j, :]), warps run over the nonzeros of the column slice their CTA owns
// (samples are range-partitioned over the cluster's CTAs exactly as in pcd.cu), the per-CTA
// partial (g_s, h_s) vectors are exchanged all-to-all through distributed shared memory, and the
// row prox + regularizer cache chain is evaluated redundantly by every warp.  The reference's
// dA[n, m, k] stash (pbcd.py:60-67) is recomputed instead of stored.
#include "common.cuh"
#include "cluster.cuh"
#include "sparsepoly_b200.h"
#include "pbcd_common.cuh"

int sp_rows_precompute_all(const sp_dataset *ds, const double *P_dk, int k, int degree, double *A,
                           cudaStream_t st);
int sp_launch_reg_cache(int mode, int degree, int d, const double *v, double *regstate, cudaStream_t st);
int sp_pbcd_wsweep(const sp_dataset *ds, const sp_wplan *wp, const int32_t *idx_feat, double *P_dk, int k,
                   const double *lams, int degree, double beta, double gamma, double eta, int reg, int loss,
                   double *yrec, double *A, double *norms, double *regstate, double *viol, cudaStream_t st);

namespace {

constexpr int PB_MAX_THREADS = 256;
constexpr int PB_MAX_CTAS = 16;

struct BlockArgs {
    int d, C, k;
    const int32_t *pos_ptr, *flag_idx, *idx_feat;
    const double *data;
    double *P;            // [d,k]
    const double *lams;   // [k]
    double beta, gamma, eta;
    int reg, loss;
    double *yrec;         // [n,2]
    double *A;            // [n,(m-1),k] or [n,k]
    double *norms;        // [d]
    double *regstate;     // cache[0..]
    double *viol;
};

// row norms: norms[j] = ||P[j,:]||_2   (squaredl21.py:37, omegacs.py:65)
__global__ void row_norms_kernel(int d, int k, const double *__restrict__ P, double *norms) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < d; j += n_warps) {
        double acc = 0.0;
        for (int s = lane; s < k; s += 32) { const double v = P[(size_t)j * k + s]; acc += v * v; }
        acc = sp_warp_allsum(acc);
        if (lane == 0) norms[j] = sqrt(acc);
    }
}

template <int KIND, int DEG, int KCH>
__global__ void __launch_bounds__(PB_MAX_THREADS) pbcd_sweep_kernel(const BlockArgs a) {
    constexpr int ND = (KIND == PK_FM) ? DEG : 1;          // dA chain length
    constexpr int NA = (KIND == PK_FM) ? DEG - 1 : 1;      // A rows stored per sample
    constexpr int NC = (KIND == PK_FM) ? DEG + 1 : 1;
    constexpr int UB = (KCH * NA <= 2) ? 4 : ((KCH * NA <= 8) ? 2 : 1);   // nonzeros in flight per warp

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar[2];

    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const int C = a.C, PS = C + 1, d = a.d, k = a.k;
    const int c = (C > 1) ? (int)cluster_ctarank() : 0;
    // dynamic smem: red[W][k] double2, then mbox[2][C][k] double2
    double2 *red = reinterpret_cast<double2 *>(smem_raw);
    double2 *mbox = red + (size_t)PB_MAX_THREADS / 32 * k;

    const double mu = sp_mu_rt(a.loss);
    const double beta = a.beta, gamma = a.gamma, eta = a.eta;
    const int reg = a.reg, loss = a.loss;
    const size_t strideA = (size_t)NA * k;

    double lam[KCH];
#pragma unroll
    for (int q = 0; q < KCH; q++) { const int s = lane + 32 * q; lam[q] = s < k ? a.lams[s] : 0.0; }

    double viol = *a.viol;
    double cache[NC];
#pragma unroll
    for (int t = 0; t < NC; t++) cache[t] = a.regstate[t];

    if (C > 1) {
        if (tid == 0) {
            mbar_init(smem_u32(&mbar[0]), 1);
            mbar_init(smem_u32(&mbar[1]), 1);
            fence_mbar_init();
        }
        __syncthreads();
        cluster_sync_all();
    }

    const bool use_norms = (reg == SP_REG_SQL21 || reg == SP_REG_OMEGACS);
    int s1 = 0, e1 = 0, j1 = 0, s2 = 0, e2 = 0, j2 = 0;
    double pnext[KCH], norm_next = 0.0;
#pragma unroll
    for (int q = 0; q < KCH; q++) pnext[q] = 0.0;
    if (0 < d) {
        s1 = a.pos_ptr[c]; e1 = a.pos_ptr[c + 1]; j1 = a.idx_feat[0];
#pragma unroll
        for (int q = 0; q < KCH; q++) { const int s = lane + 32 * q; if (s < k) pnext[q] = a.P[(size_t)j1 * k + s]; }
        if (use_norms) norm_next = a.norms[j1];
    }
    if (1 < d) { s2 = a.pos_ptr[PS + c]; e2 = a.pos_ptr[PS + c + 1]; j2 = a.idx_feat[1]; }

    for (int t = 0; t < d; t++) {
        const int s0 = s1, e0 = e1, j0 = j1;
        double pold[KCH], g[KCH], h[KCH];
#pragma unroll
        for (int q = 0; q < KCH; q++) { pold[q] = pnext[q]; g[q] = 0.0; h[q] = 0.0; }
        const double norm_old = norm_next;
        s1 = s2; e1 = e2; j1 = j2;
        if (t + 1 < d) {       // P row / norm of the next position (never touched by this step)
#pragma unroll
            for (int q = 0; q < KCH; q++) { const int s = lane + 32 * q; if (s < k) pnext[q] = a.P[(size_t)j1 * k + s]; }
            if (use_norms) norm_next = a.norms[j1];
        }
        if (t + 2 < d) {
            s2 = a.pos_ptr[(size_t)(t + 2) * PS + c];
            e2 = a.pos_ptr[(size_t)(t + 2) * PS + c + 1];
            j2 = a.idx_feat[t + 2];
        }
        if (C > 1 && tid == 0) mbar_arrive_expect_tx(smem_u32(&mbar[t & 1]), 16u * (uint32_t)C * (uint32_t)k);
        // this warp's nonzeros of the slice: e = s0 + warp + W*q, q < nq
        const int cnt = e0 - s0;
        const int nq = cnt > warp ? (cnt - warp + W - 1) / W : 0;

        // ------------------------------------------------ partial gradient / curvature sums
        for (int qb = 0; qb < nq; qb += 32) {
            const int ql = qb + lane;
            int idx_l = 0;
            double x_l = 0.0;
            if (ql < nq) {
                const int e = s0 + warp + W * ql;
                idx_l = a.flag_idx[e] & SP_ROW_MASK;
                x_l = a.data[e];
            }
            const int nloc = min(32, nq - qb);
            for (int q0 = 0; q0 < nloc; q0 += UB) {
                int iv[UB];
                double xv[UB], Av[UB][KCH][NA];
                double2 yy[UB];
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    iv[u] = __shfl_sync(0xffffffffu, idx_l, (q0 + u) & 31);
                    xv[u] = sp_shfl(x_l, (q0 + u) & 31);
                }
#pragma unroll
                for (int u = 0; u < UB; u++) {            // UB independent row gathers in flight
                    if (q0 + u < nloc) {
                        yy[u] = *reinterpret_cast<const double2 *>(a.yrec + (size_t)iv[u] * 2);
                        const double *Ai = a.A + (size_t)iv[u] * strideA;
#pragma unroll
                        for (int q = 0; q < KCH; q++) {
                            const int s = lane + 32 * q;
#pragma unroll
                            for (int r = 0; r < NA; r++) Av[u][q][r] = (s < k) ? Ai[(size_t)r * k + s] : 0.0;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    if (q0 + u < nloc) {
                        const double x = xv[u];
                        const double dl = sp_dloss_rt(loss, yy[u].x, yy[u].y);
#pragma unroll
                        for (int q = 0; q < KCH; q++) {
                            double last;
                            if (KIND == PK_FM) {
                                double dprev = x;                          // pbcd.py:9-15
#pragma unroll
                                for (int r = 1; r < ND; r++) dprev = x * (Av[u][q][r - 1] - pold[q] * dprev);
                                last = dprev;
                            } else {
                                last = x * Av[u][q][0] / (1.0 + x * pold[q]);   // pbcd_all.py:51
                            }
                            if (lane + 32 * q < k) {
                                g[q] += dl * last;                          // pbcd.py:65-67
                                h[q] += last * last;
                            }
                        }
                    }
                }
            }
        }
        if (W > 1) {
#pragma unroll
            for (int q = 0; q < KCH; q++) {
                const int s = lane + 32 * q;
                if (s < k) red[(size_t)warp * k + s] = make_double2(g[q], h[q]);
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < KCH; q++) {
                const int s = lane + 32 * q;
                g[q] = 0.0; h[q] = 0.0;
                if (s < k)
                    for (int w = 0; w < W; w++) { const double2 v = red[(size_t)w * k + s]; g[q] += v.x; h[q] += v.y; }
            }
        }
        if (C > 1) {
            const int par = t & 1;
            double2 *box = mbox + (size_t)par * C * k;
            for (int r = warp; r < C; r += W) {             // warp r-th pushes this CTA's vector to CTA r
                const uint32_t rbar = mapa_u32(smem_u32(&mbar[par]), (uint32_t)r);
#pragma unroll
                for (int q = 0; q < KCH; q++) {
                    const int s = lane + 32 * q;
                    if (s < k)
                        st_async_2f64(mapa_u32(smem_u32(&box[(size_t)c * k + s]), (uint32_t)r), g[q], h[q], rbar);
                }
            }
            mbar_wait(smem_u32(&mbar[par]), (uint32_t)((t >> 1) & 1));
#pragma unroll
            for (int q = 0; q < KCH; q++) {
                const int s = lane + 32 * q;
                g[q] = 0.0; h[q] = 0.0;
                if (s < k)
                    for (int r = 0; r < C; r++) { const double2 v = box[(size_t)r * k + s]; g[q] += v.x; h[q] += v.y; }
            }
        }

        // ------------------------------------------------ block step (redundant in every warp)
        double inv = 0.0;
#pragma unroll
        for (int q = 0; q < KCH; q++) inv += h[q];
        inv = sp_warp_allsum(inv);                            // pbcd.py:68-72
        inv = inv * mu;
        inv = inv + beta;
        double pnew[KCH];
#pragma unroll
        for (int q = 0; q < KCH; q++) {
            double gr = g[q] * lam[q];                        // pbcd.py:74-78
            gr = gr + beta * pold[q];
            gr = gr / inv;
            pnew[q] = pold[q] - eta * gr;
        }
        double strength = eta * gamma / inv;
        // ---- prox_bcd (l1.py:44-45, l21.py:33-38, squaredl21.py:45-55, omegacs.py:83-106)
        if (reg == SP_REG_L1) {
#pragma unroll
            for (int q = 0; q < KCH; q++) pnew[q] = sp_soft_threshold(pnew[q], strength);
        } else {
            if (reg == SP_REG_SQL21) {
                const double den = 1.0 + 2.0 * strength;
#pragma unroll
                for (int q = 0; q < KCH; q++) pnew[q] = pnew[q] / den;
            }
            double dot = 0.0;
#pragma unroll
            for (int q = 0; q < KCH; q++) dot += pnew[q] * pnew[q];
            dot = sp_warp_allsum(dot);
            const double l2 = sqrt(dot);
            double dc[NC + 1];
#pragma unroll
            for (int u = 0; u <= NC; u++) dc[u] = 0.0;
            double norm_j = norm_old;
            if (reg == SP_REG_SQL21) {
                if (cache[0] < norm_j) cache[0] = warp_sum_array(a.norms, d);
                const double dcache = cache[0] - norm_j;
                strength = 2.0 * dcache * strength / (1.0 + 2.0 * strength);
            } else if (reg == SP_REG_OMEGACS) {
                if (KIND == PK_FM) {
                    dc[1] = 1.0;
                    bool neg = false;
#pragma unroll
                    for (int deg = 2; deg <= DEG; deg++) {
                        double v = cache[deg - 1];
                        v = v - dc[deg - 1] * norm_j;
                        dc[deg] = v;
                        neg = neg || (v < 0.0);
                    }
                    if (neg) {                                // omegacs.py:90-96 recovery branch
                        norm_j = 0.0;
                        double ec[DEG + 1];
                        warp_esp<DEG>(a.norms, d, j0, DEG - 1, ec);
#pragma unroll
                        for (int u = 0; u <= DEG; u++) cache[u] = ec[u];
                        dc[0] = 0.0; dc[1] = 1.0;
#pragma unroll
                        for (int deg = 2; deg <= DEG; deg++) dc[deg] = cache[DEG - 1];
                    }
                    strength = strength * dc[DEG];
                } else {
                    cache[0] = cache[0] / (1.0 + norm_j);
                    strength = strength * cache[0];
                }
            }
            if (l2 > strength) {
                const double sc = 1.0 - strength / l2;
#pragma unroll
                for (int q = 0; q < KCH; q++) pnew[q] = pnew[q] * sc;
            } else {
#pragma unroll
                for (int q = 0; q < KCH; q++) pnew[q] = 0.0;
            }
            // ---- update_cache_pbcd (squaredl21.py:40-43, omegacs.py:68-81)
            if (reg == SP_REG_SQL21 || reg == SP_REG_OMEGACS) {
                double dn = 0.0;
#pragma unroll
                for (int q = 0; q < KCH; q++) dn += pnew[q] * pnew[q];
                dn = sp_warp_allsum(dn);
                const double l2n = sqrt(dn);
                if (reg == SP_REG_SQL21) {
                    cache[0] = cache[0] - norm_j;
                    cache[0] = cache[0] + l2n;
                } else if (KIND == PK_FM) {
                    bool neg = false;
#pragma unroll
                    for (int deg = 1; deg <= DEG; deg++) {
                        cache[deg] = cache[deg] + dc[deg] * l2n;
                        cache[deg] = cache[deg] - dc[deg] * norm_j;
                    }
#pragma unroll
                    for (int deg = 0; deg <= DEG; deg++) neg = neg || (cache[deg] < 0.0);
                    if (neg) {                                // omegacs.py:75-76: full recompute
                        if (c == 0 && tid == 0) a.norms[j0] = l2n;
                        // every warp recomputes from global norms with entry j0 := l2n
                        double ec[DEG + 1], e2[DEG + 1];
                        warp_esp<DEG>(a.norms, d, j0, DEG, ec);
                        e2[0] = ec[0];
#pragma unroll
                        for (int u = 1; u <= DEG; u++) e2[u] = ec[u] + ec[u - 1] * l2n;
#pragma unroll
                        for (int u = 0; u <= DEG; u++) cache[u] = e2[u];
                    }
                } else {
                    cache[0] = cache[0] * (1.0 + l2n);
                }
                if (c == 0 && tid == 0) a.norms[j0] = l2n;
            }
        }
        double upd[KCH], l1 = 0.0;
        bool moved = false;
#pragma unroll
        for (int q = 0; q < KCH; q++) {
            upd[q] = pold[q] - pnew[q];
            l1 += fabs(upd[q]);
            moved = moved || (upd[q] != 0.0);
        }
        viol += sp_warp_allsum(l1);                           // pbcd.py:146 norm(updates, 1)
        moved = __any_sync(0xffffffffu, moved);
        if (c == 0 && warp == 0) {
#pragma unroll
            for (int q = 0; q < KCH; q++) {
                const int s = lane + 32 * q;
                if (s < k) a.P[(size_t)j0 * k + s] = pnew[q];
            }
        }

        // ------------------------------------------------ synchronize predictions and caches
        if (KIND == PK_ALL || moved) {
            for (int qb = 0; qb < nq; qb += 32) {
                const int ql = qb + lane;
                int idx_l = 0;
                double x_l = 0.0;
                if (ql < nq) {
                    const int e = s0 + warp + W * ql;
                    idx_l = a.flag_idx[e] & SP_ROW_MASK;
                    x_l = a.data[e];
                }
                const int nloc = min(32, nq - qb);
                for (int q0 = 0; q0 < nloc; q0 += UB) {
                    int iv[UB];
                    double xv[UB], Av[UB][KCH][NA], ypv[UB];
#pragma unroll
                    for (int u = 0; u < UB; u++) {
                        iv[u] = __shfl_sync(0xffffffffu, idx_l, (q0 + u) & 31);
                        xv[u] = sp_shfl(x_l, (q0 + u) & 31);
                    }
#pragma unroll
                    for (int u = 0; u < UB; u++) {
                        if (q0 + u < nloc) {
                            ypv[u] = a.yrec[(size_t)iv[u] * 2];
                            const double *Ai = a.A + (size_t)iv[u] * strideA;
#pragma unroll
                            for (int q = 0; q < KCH; q++) {
                                const int s = lane + 32 * q;
#pragma unroll
                                for (int r = 0; r < NA; r++) Av[u][q][r] = (s < k) ? Ai[(size_t)r * k + s] : 0.0;
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UB; u++) {
                        if (q0 + u < nloc) {
                            const double x = xv[u];
                            double *Ai = a.A + (size_t)iv[u] * strideA;
                            double dy = 0.0, dy2 = 0.0;
#pragma unroll
                            for (int q = 0; q < KCH; q++) {
                                const int s = lane + 32 * q;
                                if (s < k) {
                                    if (KIND == PK_FM) {
                                        double dprev = x;                  // pbcd.py:138-144
#pragma unroll
                                        for (int r = 1; r < ND; r++) {
                                            const double Aold = Av[u][q][r - 1];
                                            const double dcur = x * (Aold - pold[q] * dprev);
                                            Ai[(size_t)(r - 1) * k + s] = Aold - upd[q] * dprev;
                                            dprev = dcur;
                                        }
                                        dy += (lam[q] * upd[q]) * dprev;
                                    } else {
                                        double Aval = Av[u][q][0];         // pbcd_all.py:121-126
                                        dy += lam[q] * Aval;
                                        Aval = Aval / (1.0 + x * pold[q]);
                                        Aval = Aval * (1.0 + x * pnew[q]);
                                        Ai[s] = Aval;
                                        dy2 += lam[q] * Aval;
                                    }
                                }
                            }
                            dy = sp_warp_allsum(dy);
                            if (KIND == PK_ALL) dy2 = sp_warp_allsum(dy2);
                            if (lane == 0) {
                                double yp = ypv[u] - dy;
                                if (KIND == PK_ALL) yp = yp + dy2;
                                a.yrec[(size_t)iv[u] * 2] = yp;
                            }
                        }
                    }
                }
            }
        }
        if (W > 1) __syncthreads(); else __syncwarp();
    }

    if (c == 0 && tid == 0) {
        *a.viol = viol;
#pragma unroll
        for (int t = 0; t < NC; t++) a.regstate[t] = cache[t];
    }
    if (C > 1) cluster_sync_all();
}

template <int KIND, int DEG, int KCH>
int launch_block(const BlockArgs &a, int threads, cudaStream_t st) {
    auto kern = pbcd_sweep_kernel<KIND, DEG, KCH>;
    const size_t smem = ((size_t)PB_MAX_THREADS / 32 + 2 * (size_t)a.C) * a.k * sizeof(double2);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return sp_check_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    if (a.C > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return sp_check_cuda(e, "cudaFuncSetAttribute(NonPortableClusterSizeAllowed)");
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a.C, 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    sp_prof_begin(SP_PROF_SWEEP_PBCD, st);
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, a);
    sp_prof_end(st);
    return sp_check_cuda(le, "pbcd_sweep_kernel launch");
}

template <int KIND, int DEG>
int dispatch_kch(const BlockArgs &a, int threads, cudaStream_t st) {
    if (a.k <= 32) return launch_block<KIND, DEG, 1>(a, threads, st);
    if (a.k <= 64) return launch_block<KIND, DEG, 2>(a, threads, st);
    if (a.k <= 128) return launch_block<KIND, DEG, 4>(a, threads, st);
    sp_set_error("pbcd: n_components=%d > 128 is not supported by the CUDA backend", a.k);
    return SP_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" int sp_pbcd_epoch(const sp_dataset *ds, const sp_plan *plan, double *P_dk, int k,
                             const double *lams, int degree, double beta, double gamma, double eta,
                             int reg, int loss, double *yrec, double *A, double *reg_norms,
                             double *regstate, double *viol, sp_stream stream) {
    if (!ds || !plan || !ds->csc_data || !ds->csr_indptr || !plan->idx_feat || !P_dk || !lams || !yrec || !A ||
        !reg_norms || !regstate || !viol || k <= 0 || (!plan->win && (!plan->pos_ptr || !plan->flag_idx))) {
        sp_set_error("sp_pbcd_epoch: invalid argument");
        return SP_ERR_INVALID;
    }
    if (!plan->win && (plan->n_cta < 1 || plan->n_cta > PB_MAX_CTAS || (plan->n_cta & (plan->n_cta - 1)) ||
        plan->threads < 32 || plan->threads > PB_MAX_THREADS || plan->threads % 32)) {
        sp_set_error("sp_pbcd_epoch: bad plan (n_cta power of two <= %d, threads multiple of 32 <= %d)",
                     PB_MAX_CTAS, PB_MAX_THREADS);
        return SP_ERR_INVALID;
    }
    if (reg != SP_REG_L1 && reg != SP_REG_L21 && reg != SP_REG_SQL21 && reg != SP_REG_OMEGACS) {
        sp_set_error("regularizer id %d does not implement the pbcd protocol (use l1, l21, squaredl21 or omegacs)", reg);
        return SP_ERR_UNSUPPORTED;
    }
    if (reg == SP_REG_SQL21 && degree != 2) {
        sp_set_error(degree == -1 ? "squaredl21 is not available for all-subsets models"
                                  : "SquaredL21 supports only degree=2.");
        return SP_ERR_UNSUPPORTED;
    }
    if (!(degree == -1 || (degree >= 2 && degree <= SP_MAXDEG))) {
        sp_set_error("pbcd degree %d is not supported by the CUDA backend (2..%d or -1)", degree, SP_MAXDEG);
        return SP_ERR_UNSUPPORTED;
    }
    if (loss < 0 || loss > 2) { sp_set_error("unknown loss id %d", loss); return SP_ERR_INVALID; }
    const int d = ds->n_features;
    if (d == 0) return SP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = sp_rows_precompute_all(ds, P_dk, k, degree, A, st);               // pbcd.py:107
    if (rc) return rc;
    if (reg == SP_REG_SQL21 || reg == SP_REG_OMEGACS) {                         // pbcd.py:109
        int blocks = (d + 7) / 8; if (blocks > 148 * 8) blocks = 148 * 8;
        sp_prof_begin(SP_PROF_REGCACHE, st);
        row_norms_kernel<<<blocks, 256, 0, st>>>(d, k, P_dk, reg_norms);
        sp_prof_end(st);
        SP_LAUNCH_CHECK("row_norms_kernel");
        const int mode = (reg == SP_REG_SQL21) ? 0 : (degree == -1 ? 2 : 1);
        rc = sp_launch_reg_cache(mode, degree, d, reg_norms, regstate, st);
        if (rc) return rc;
    }
    if (plan->win)
        return sp_pbcd_wsweep(ds, plan->win, plan->idx_feat, P_dk, k, lams, degree, beta, gamma, eta, reg, loss, yrec,
                              A, reg_norms, regstate, viol, st);
    BlockArgs a = {};
    a.d = d; a.C = plan->n_cta; a.k = k;
    a.pos_ptr = plan->pos_ptr; a.flag_idx = plan->flag_idx; a.idx_feat = plan->idx_feat;
    a.data = ds->csc_data; a.P = P_dk; a.lams = lams;
    a.beta = beta; a.gamma = gamma; a.eta = eta; a.reg = reg; a.loss = loss;
    a.yrec = yrec; a.A = A; a.norms = reg_norms; a.regstate = regstate; a.viol = viol;
    switch (degree) {
    case -1: return dispatch_kch<PK_ALL, 1>(a, plan->threads, st);
    case 2: return dispatch_kch<PK_FM, 2>(a, plan->threads, st);
    case 3: return dispatch_kch<PK_FM, 3>(a, plan->threads, st);
    case 4: return dispatch_kch<PK_FM, 4>(a, plan->threads, st);
    case 5: return dispatch_kch<PK_FM, 5>(a, plan->threads, st);
    }
    return SP_ERR_UNSUPPORTED;
}
