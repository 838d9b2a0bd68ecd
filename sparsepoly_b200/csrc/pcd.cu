// Sequential coordinate sweeps on ONE thread-block cluster (sm_100a):
//   * cd_linear._cd_linear_epoch                     (reference optimizer/cd_linear.py:8-33)
//   * pcd.pcd_epoch  (_update + synchronize loop)    (reference optimizer/pcd.py:33-137)
//   * pcd_all.pcd_epoch                              (reference optimizer/pcd_all.py:21-102)
//
// Design (see DESIGN.md "pcd sweep"):  coordinate order is a true dependency chain, so a sweep
// is latency-bound, not bandwidth-bound.  One cluster of C CTAs walks the d coordinates in order.
// Samples are range-partitioned over the CTAs (CTA c owns rows [c*chunk,(c+1)*chunk)), so a
// sample's record {y_pred, y, A^1..A^{m-1}} is only ever touched by one SM: no inter-SM memory
// hazards, only an all-to-all exchange of the per-CTA partial sums (g_c, h_c) through
// distributed shared memory (st.async + mbarrier complete_tx), one hop (~DSMEM latency) per
// coordinate.  Every thread then runs the ~20-flop scalar chain (step, prox, regularizer
// cache) redundantly, so no broadcast is needed.  Column slices, records and P entries of the
// next positions are software-pipelined through registers (distance 1-3 positions); records
// of samples that the previous position is still rewriting are tagged by the plan
// (SP_FLAG_BIT in flag_idx) and re-read after the end-of-step barrier.
#include <stdarg.h>

#include "common.cuh"
#include "cluster.cuh"
#include "sparsepoly_b200.h"

int sp_rows_precompute_one(const sp_dataset *ds, const double *p_s, int degree, double *rec,
                           int rec_stride, cudaStream_t st);
int sp_launch_reg_cache(int mode, int degree, int d, const double *v, double *regstate, cudaStream_t st);

namespace {

enum { KIND_LINEAR = 0, KIND_FM = 1, KIND_ALL = 2 };
constexpr int SWEEP_MAX_THREADS = 256;
constexpr int SWEEP_MAX_CTAS = 16;

struct SweepArgs {
    int d, C;
    const int32_t *pos_ptr;    // [d*(C+1)]
    const int32_t *flag_idx;   // [nnz]
    const double *data;        // [nnz] CSC values
    const int32_t *idx_feat;   // [d]
    double *prow;              // P[s, :] (or w)
    const double *cns;         // col_norm_sq (linear only)
    const double *lam_ptr;     // &lams[s] (FM / all-subsets)
    double ab;                 // alpha (linear) or beta
    double gamma, eta;
    int reg;
    double *rec;
    int stride;
    double *regstate;          // in/out regularizer scalars
    double *viol;              // in/out running sum of |updates|
};

// ------------------------------------------------------------------------- record access
template <int R> __device__ __forceinline__ void load_rec(const double *p, double (&r)[R]) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
    for (int u = 0; u < R / 2; u++) { double2 v = q[u]; r[2 * u] = v.x; r[2 * u + 1] = v.y; }
    if (R & 1) r[R - 1] = p[R - 1];
}

// per-nonzero gradient terms.  r = {y_pred, y, A^1.. }.  dA[] keeps the chain for the scatter.
template <int KIND, int DEG, int LOSS, int R, int ND>
__device__ __forceinline__ void nz_terms(const double (&r)[R], double x, double pold, double (&dA)[ND],
                                         double &tg, double &th) {
    const double dl = sp_dloss<LOSS>(r[0], r[1]);
    if (KIND == KIND_LINEAR) {
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
__device__ __forceinline__ void nz_scatter(double *p, double (&r)[R], const double (&dA)[ND], double x,
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

template <int KIND, int DEG, int LOSS>
__global__ void __launch_bounds__(SWEEP_MAX_THREADS) sweep_kernel(const SweepArgs a) {
    constexpr int NA = (KIND == KIND_FM) ? DEG - 1 : (KIND == KIND_ALL ? 1 : 0);
    constexpr int R = 2 + NA;
    constexpr int ND = (KIND == KIND_FM) ? DEG : 1;
    constexpr int NC = (KIND == KIND_FM) ? DEG : 1;          // regularizer cache scalars

    __shared__ double2 red[SWEEP_MAX_THREADS / 32];
    __shared__ __align__(16) double2 mbox[2][SWEEP_MAX_CTAS];
    __shared__ __align__(8) unsigned long long mbar[2];

    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const int C = a.C, PS = C + 1, d = a.d;
    const int c = (C > 1) ? (int)cluster_ctarank() : 0;
    const int stride = a.stride;
    const double mu = sp_mu<LOSS>();
    const double lam = (KIND == KIND_LINEAR) ? 1.0 : *a.lam_ptr;
    const double ab = a.ab, gamma = a.gamma, eta = a.eta;
    const int reg = a.reg;

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

    // ---- pipeline registers: position t (0), t+1 (1), t+2 (2), t+3 (3)
    int s0 = 0, e0 = 0, j0 = 0, s1 = 0, e1 = 0, j1 = 0, s2 = 0, e2 = 0, j2 = 0, s3 = 0, e3 = 0, j3 = 0;
    int fi0 = 0, fi1 = 0, fi2 = 0;
    double x0 = 0.0, x1 = 0.0, x2 = 0.0, pold0 = 0.0, pold1 = 0.0, cn0 = 0.0, cn1 = 0.0;
    double r0[R], r1[R];
#pragma unroll
    for (int u = 0; u < R; u++) { r0[u] = 0.0; r1[u] = 0.0; }

    if (0 < d) { s0 = a.pos_ptr[c]; e0 = a.pos_ptr[c + 1]; j0 = a.idx_feat[0]; }
    if (1 < d) { s1 = a.pos_ptr[PS + c]; e1 = a.pos_ptr[PS + c + 1]; j1 = a.idx_feat[1]; }
    if (2 < d) { s2 = a.pos_ptr[2 * PS + c]; e2 = a.pos_ptr[2 * PS + c + 1]; j2 = a.idx_feat[2]; }
    if (s0 + tid < e0) {
        fi0 = a.flag_idx[s0 + tid];
        x0 = a.data[s0 + tid];
        load_rec<R>(a.rec + (size_t)(fi0 & 0x7fffffff) * stride, r0);
    }
    if (s1 + tid < e1) { fi1 = a.flag_idx[s1 + tid]; x1 = a.data[s1 + tid]; }
    if (0 < d) { pold0 = a.prow[j0]; if (KIND == KIND_LINEAR) cn0 = a.cns[j0]; }

    for (int t = 0; t < d; t++) {
        // ------------------------------------------------ issue the loads of future positions
        if (t + 3 < d) {
            s3 = a.pos_ptr[(size_t)(t + 3) * PS + c];
            e3 = a.pos_ptr[(size_t)(t + 3) * PS + c + 1];
            j3 = a.idx_feat[t + 3];
        } else { s3 = 0; e3 = 0; j3 = 0; }
        if (s2 + tid < e2) { fi2 = a.flag_idx[s2 + tid]; x2 = a.data[s2 + tid]; }
        const bool has1 = s1 + tid < e1;
        if (has1 && fi1 >= 0) load_rec<R>(a.rec + (size_t)fi1 * stride, r1);
        if (t + 1 < d) { pold1 = a.prow[j1]; if (KIND == KIND_LINEAR) cn1 = a.cns[j1]; }
        if (C > 1 && tid == 0) mbar_arrive_expect_tx(smem_u32(&mbar[t & 1]), 16u * (uint32_t)C);

        // ------------------------------------------------ gradient / curvature partial sums
        const bool has0 = s0 + tid < e0;
        const int i0 = fi0 & 0x7fffffff;
        double tg = 0.0, th = 0.0;
        double dA0[ND];
#pragma unroll
        for (int u = 0; u < ND; u++) dA0[u] = 0.0;
        if (has0) {
            if (fi0 < 0) load_rec<R>(a.rec + (size_t)i0 * stride, r0);   // hazard: re-read
            nz_terms<KIND, DEG, LOSS, R, ND>(r0, x0, pold0, dA0, tg, th);
        }
        for (int e = s0 + T + tid; e < e0; e += T) {          // slices longer than the CTA
            double rr[R], dd[ND];
            load_rec<R>(a.rec + (size_t)(a.flag_idx[e] & 0x7fffffff) * stride, rr);
            nz_terms<KIND, DEG, LOSS, R, ND>(rr, a.data[e], pold0, dd, tg, th);
        }
        tg = sp_warp_allsum(tg);
        if (KIND != KIND_LINEAR) th = sp_warp_allsum(th);
        if (W > 1) {
            if (lane == 0) red[warp] = make_double2(tg, th);
            __syncthreads();
            tg = 0.0; th = 0.0;
            for (int w = 0; w < W; w++) { const double2 v = red[w]; tg += v.x; th += v.y; }
        }
        if (C > 1) {
            const int par = t & 1;
            if (tid < C)
                st_async_2f64(mapa_u32(smem_u32(&mbox[par][c]), (uint32_t)tid), tg, th,
                              mapa_u32(smem_u32(&mbar[par]), (uint32_t)tid));
            mbar_wait(smem_u32(&mbar[par]), (uint32_t)((t >> 1) & 1));
            tg = 0.0; th = 0.0;
            for (int r = 0; r < C; r++) { const double2 v = mbox[par][r]; tg += v.x; th += v.y; }
        }

        // ------------------------------------------------ scalar chain (redundant in every thread)
        double pnew, upd;
        if (KIND == KIND_LINEAR) {
            double u = tg + ab * pold0;                       // cd_linear.py:19-22
            const double inv = mu * cn0 + ab;
            u = u / inv;
            pnew = pold0 - u;
            upd = u;
        } else {
            double inv = th * mu;                             // pcd.py:59-68 / pcd_all.py:34-41
            inv = inv + ab;
            double u = tg * lam;
            u = u + ab * pold0;
            u = u / inv;
            double p = pold0 - eta * u;
            double strength = eta * gamma / inv;
            const double a_old = fabs(pold0);
            if (reg == SP_REG_L1) {                           // l1.py:32-33
                pnew = sp_soft_threshold(p, strength);
            } else if (reg == SP_REG_SQL12) {                 // squaredl12.py:52-57, :47-50
                const double dcache = cache[0] - a_old;
                p = p / (1.0 + 2.0 * strength);
                const double sign = p > 0.0 ? 1.0 : -1.0;
                double m = fabs(p) - 2.0 * strength * dcache / (1.0 + 2.0 * strength);
                if (!(m > 0.0)) m = 0.0;
                pnew = sign * m;
                cache[0] = cache[0] - a_old;
                cache[0] = cache[0] + fabs(pnew);
            } else {                                          // omegati.py:82-104
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
            upd = pold0 - pnew;
        }
        viol += fabs(upd);
        if (c == 0 && tid == 0) a.prow[j0] = pnew;

        // ------------------------------------------------ synchronize predictions and caches
        if (KIND == KIND_ALL || upd != 0.0) {
            if (has0) nz_scatter<KIND, DEG, R, ND>(a.rec + (size_t)i0 * stride, r0, dA0, x0, lam, upd, pold0, pnew);
            for (int e = s0 + T + tid; e < e0; e += T) {
                double rr[R], dd[ND];
                double *p = a.rec + (size_t)(a.flag_idx[e] & 0x7fffffff) * stride;
                const double x = a.data[e];
                load_rec<R>(p, rr);
                if (KIND == KIND_FM) {
                    dd[0] = x;
#pragma unroll
                    for (int u = 1; u < ND; u++) dd[u] = x * (rr[1 + u] - pold0 * dd[u - 1]);
                } else {
                    dd[0] = 0.0;
                }
                nz_scatter<KIND, DEG, R, ND>(p, rr, dd, x, lam, upd, pold0, pnew);
            }
        }
        if (W > 1) __syncthreads(); else __syncwarp();

        // ------------------------------------------------ rotate the pipeline
        s0 = s1; e0 = e1; j0 = j1; pold0 = pold1; cn0 = cn1; fi0 = fi1; x0 = x1;
#pragma unroll
        for (int u = 0; u < R; u++) r0[u] = r1[u];
        s1 = s2; e1 = e2; j1 = j2; fi1 = fi2; x1 = x2;
        s2 = s3; e2 = e3; j2 = j3;
    }

    if (c == 0 && tid == 0) {
        *a.viol = viol;
        if (KIND != KIND_LINEAR) {
#pragma unroll
            for (int t = 0; t < NC; t++) a.regstate[t] = cache[t];
        }
    }
    if (C > 1) cluster_sync_all();   // no CTA may exit while peers can still write its smem
}

template <int KIND, int DEG, int LOSS>
int launch_sweep(const SweepArgs &a, int threads, cudaStream_t st) {
    auto kern = sweep_kernel<KIND, DEG, LOSS>;
    if (a.C > 8) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return sp_check_cuda(e, "cudaFuncSetAttribute(NonPortableClusterSizeAllowed)");
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a.C, 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    sp_prof_begin(SP_PROF_SWEEP_PCD, st);
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, a);
    sp_prof_end(st);
    return sp_check_cuda(le, "sweep_kernel launch");
}

template <int KIND, int DEG>
int dispatch_loss(int loss, const SweepArgs &a, int threads, cudaStream_t st) {
    switch (loss) {
    case SP_LOSS_SQUARED: return launch_sweep<KIND, DEG, SP_LOSS_SQUARED>(a, threads, st);
    case SP_LOSS_LOGISTIC: return launch_sweep<KIND, DEG, SP_LOSS_LOGISTIC>(a, threads, st);
    case SP_LOSS_SQHINGE: return launch_sweep<KIND, DEG, SP_LOSS_SQHINGE>(a, threads, st);
    default: sp_set_error("unknown loss id %d", loss); return SP_ERR_INVALID;
    }
}

int check_plan(const sp_dataset *ds, const sp_plan *plan, const char *who) {
    if (!ds || !plan || !ds->csc_data || !plan->pos_ptr || !plan->flag_idx || !plan->idx_feat) {
        sp_set_error("%s: dataset/plan pointers missing", who);
        return SP_ERR_INVALID;
    }
    if (plan->n_cta < 1 || plan->n_cta > SWEEP_MAX_CTAS || (plan->n_cta & (plan->n_cta - 1))) {
        sp_set_error("%s: plan.n_cta must be a power of two in 1..%d", who, SWEEP_MAX_CTAS);
        return SP_ERR_INVALID;
    }
    if (plan->threads < 32 || plan->threads > SWEEP_MAX_THREADS || plan->threads % 32) {
        sp_set_error("%s: plan.threads must be a multiple of 32 in 32..%d", who, SWEEP_MAX_THREADS);
        return SP_ERR_INVALID;
    }
    return SP_OK;
}

}  // namespace

int sp_min_rec_stride(int degree) {   // record = {y_pred, y, A^1..A^{m-1}}, even number of doubles
    const int r = (degree == -1) ? 3 : (degree < 1 ? 2 : degree + 1);
    return (r + 3) & ~3;
}

extern "C" int sp_rec_stride(int degree) { return sp_min_rec_stride(degree); }

extern "C" int sp_cd_linear_epoch(const sp_dataset *ds, const sp_plan *plan, double *w,
                                  const double *col_norm_sq, double alpha, int loss, double *rec,
                                  int rec_stride, double *viol, sp_stream stream) {
    int rc = check_plan(ds, plan, "sp_cd_linear_epoch");
    if (rc) return rc;
    if (!w || !col_norm_sq || !rec || !viol || rec_stride < 2 || (rec_stride & 1)) {
        sp_set_error("sp_cd_linear_epoch: invalid argument");
        return SP_ERR_INVALID;
    }
    if (ds->n_features == 0) return SP_OK;
    SweepArgs a = {};
    a.d = ds->n_features; a.C = plan->n_cta;
    a.pos_ptr = plan->pos_ptr; a.flag_idx = plan->flag_idx; a.data = ds->csc_data;
    a.idx_feat = plan->idx_feat;
    a.prow = w; a.cns = col_norm_sq; a.lam_ptr = nullptr; a.ab = alpha; a.gamma = 0.0; a.eta = 1.0;
    a.reg = SP_REG_L1; a.rec = rec; a.stride = rec_stride; a.regstate = viol; a.viol = viol;
    return dispatch_loss<KIND_LINEAR, 1>(loss, a, plan->threads, (cudaStream_t)stream);
}

extern "C" int sp_pcd_epoch(const sp_dataset *ds, const sp_plan *plan, double *P_kd, int k,
                            const double *lams, int degree, double beta, double gamma, double eta,
                            int reg, int loss, double *rec, int rec_stride, double *regstate,
                            double *viol, const int32_t *idx_comp_host, sp_stream stream) {
    int rc = check_plan(ds, plan, "sp_pcd_epoch");
    if (rc) return rc;
    if (!P_kd || !lams || !rec || !regstate || !viol || !idx_comp_host || k <= 0 || !ds->csr_indptr) {
        sp_set_error("sp_pcd_epoch: invalid argument");
        return SP_ERR_INVALID;
    }
    // solver x regularizer x degree support matrix (reference README.md:25-31; the reference
    // fails with a numba TypingError / ValueError inside the jitclass for the others)
    if (reg != SP_REG_L1 && reg != SP_REG_SQL12 && reg != SP_REG_OMEGATI) {
        sp_set_error("regularizer id %d does not implement the pcd protocol (use l1, squaredl12 or omegati)", reg);
        return SP_ERR_UNSUPPORTED;
    }
    if (reg == SP_REG_SQL12 && degree != 2) {
        sp_set_error(degree == -1 ? "squaredl12 is not available for all-subsets models"
                                  : "SquaredL12 supports only degree=2.");
        return SP_ERR_UNSUPPORTED;
    }
    if (!(degree == -1 || (degree >= 2 && degree <= SP_MAXDEG))) {
        sp_set_error("pcd degree %d is not supported by the CUDA backend (2..%d or -1)", degree, SP_MAXDEG);
        return SP_ERR_UNSUPPORTED;
    }
    if (rec_stride < sp_min_rec_stride(degree) || (rec_stride & 1)) {
        sp_set_error("sp_pcd_epoch: rec_stride %d too small for degree %d", rec_stride, degree);
        return SP_ERR_INVALID;
    }
    const int d = ds->n_features;
    if (d == 0) return SP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    for (int ss = 0; ss < k; ss++) {
        const int s = idx_comp_host[ss];
        if (s < 0 || s >= k) { sp_set_error("sp_pcd_epoch: bad component index %d", s); return SP_ERR_INVALID; }
        double *prow = P_kd + (size_t)s * d;
        rc = sp_rows_precompute_one(ds, prow, degree, rec, rec_stride, st);      // pcd.py:94 / pcd_all.py:65
        if (rc) return rc;
        if (reg != SP_REG_L1) {                                                   // pcd.py:96
            const int mode = (reg == SP_REG_SQL12) ? 0 : (degree == -1 ? 2 : 1);
            rc = sp_launch_reg_cache(mode, degree, d, prow, regstate, st);
            if (rc) return rc;
        }
        SweepArgs a = {};
        a.d = d; a.C = plan->n_cta;
        a.pos_ptr = plan->pos_ptr; a.flag_idx = plan->flag_idx; a.data = ds->csc_data;
        a.idx_feat = plan->idx_feat;
        a.prow = prow; a.cns = nullptr; a.lam_ptr = lams + s; a.ab = beta; a.gamma = gamma; a.eta = eta;
        a.reg = reg; a.rec = rec; a.stride = rec_stride; a.regstate = regstate; a.viol = viol;
        switch (degree) {
        case -1: rc = dispatch_loss<KIND_ALL, 1>(loss, a, plan->threads, st); break;
        case 2: rc = dispatch_loss<KIND_FM, 2>(loss, a, plan->threads, st); break;
        case 3: rc = dispatch_loss<KIND_FM, 3>(loss, a, plan->threads, st); break;
        case 4: rc = dispatch_loss<KIND_FM, 4>(loss, a, plan->threads, st); break;
        case 5: rc = dispatch_loss<KIND_FM, 5>(loss, a, plan->threads, st); break;
        }
        if (rc) return rc;
    }
    return SP_OK;
}
