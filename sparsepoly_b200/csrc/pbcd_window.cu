// Window sweep for the BLOCK coordinate descent (reference optimizer/pbcd.py:36-148,
// pbcd_all.py:23-132): the organisation of pcd_window.cu applied to rows P[j,:] of k <= 32 components.
//   * bulk CTAs: BASE(w) reduces the cold nonzeros of window w's columns to (g_s, h_s), s < k, per
//     position (warp per nonzero, lane = component); WB(w) applies the published row updates to
//     the cold records (A rows + y_pred);
//   * engine CTA: hot sample records {A[i, :, :], y_pred, y} live in shared-memory slots; worker
//     warps (lane = component) add the hot terms, form the shared step size (one warp all-reduce),
//     the per-component Newton step and the row norm, and hand the row to the chain warp, which
//     runs prox_bcd + the regularizer cache (l1.py:44-45, l21.py:33-38, squaredl21.py:36-55,
//     omegacs.py:52-106, incl. the negative-drift recovery branches) in coordinate order and
//     publishes the row update; workers write the hot records back.
// Hand-over counters, hot/cold split and dependency flags are those of the pcd window sweep (wplan.cu).
#include "common.cuh"
#include "cluster.cuh"
#include "pbcd_common.cuh"
#include "sparsepoly_b200.h"

namespace {

constexpr int BT = 384;                       // chain warp + 11 worker warps
constexpr int BWIN = 64;                      // most positions per window
constexpr int BW_REC_BYTES = 160 * 1024;      // slots + hot nonzeros
constexpr int BPARTS = 8;                     // most bulk CTAs sharing one column
constexpr int BRING = 4;                      // windows of cold partial sums kept (ring)
constexpr int BENT = SP_PBCD_ENT_PER_SLOT;    // hot nonzeros staged per slot of capacity
// per position: pre / result / pold rows (32 doubles each), 2 cells, norm, hp, flag, 2 mbarriers
constexpr int BW_AUX_BYTES = BWIN * (3 * 32 * 8 + 2 * 16 + 8 + 4 + 4 + 16) + 64;

struct BWArgs {
    int d, k, B, H, nwin, slot_cap, slotsz, reg, loss;
    const int32_t *indptr, *cflag;
    const double *data;
    const int32_t *idx_feat, *ht_ptr, *h_sd;
    const double *h_x;
    const int32_t *n_slots, *slot_row;
    const double *P;          // [d,k], read-only during the sweep
    const double *lams;
    double beta, gamma, eta;
    double *yrec, *A, *norms, *regstate, *viol;
    double *res;              // [d][2][k] (update, new value)
    double *base;             // [BRING][BWIN][BPARTS][2][32] cold (g, h) partial sums, ring over windows
    int *base_cnt, *wb_cnt, *eng_done;
};

struct __align__(16) BCell { double v; long long tag; };

__device__ __forceinline__ int ld_acquire_b(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_ge_b(const int *p, int v, bool sleep) {
    while (ld_acquire_b(p) < v) { if (sleep) __nanosleep(64); }
}
__device__ __forceinline__ void st_release_b(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_b(int *p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ BCell bcell_load(const BCell *c) {
    BCell r;
    unsigned long long a, b;
    asm volatile("ld.volatile.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(smem_u32(c)) : "memory");
    r.v = __longlong_as_double((long long)a);
    r.tag = (long long)b;
    return r;
}
__device__ __forceinline__ void bcell_store(BCell *c, double v, long long tag) {
    asm volatile("st.volatile.shared.v2.b64 [%0], {%1, %2};" ::"r"(smem_u32(c)), "l"(__double_as_longlong(v)),
                 "l"(tag)
                 : "memory");
}
__device__ __forceinline__ int bflag_load(const int *p) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void bflag_store(int *p, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// last element of the dA chain of one (sample, component) (pbcd.py:9-15 / pbcd_all.py:51)
template <int KIND, int DEG, int NA>
__device__ __forceinline__ double dA_last(const double (&Av)[NA], double x, double pold) {
    if (KIND == PK_FM) {
        double dprev = x;
#pragma unroll
        for (int r = 1; r < DEG; r++) dprev = x * (Av[r - 1] - pold * dprev);
        return dprev;
    }
    return x * Av[0] / (1.0 + x * pold);
}

// synchronize step of one (sample, component): updates Av in place, returns the y_pred terms
// (FM: dy = lam*upd*dA[m-1]; all-subsets: dy = lam*A_old, dy2 = lam*A_new)   pbcd.py:138-144, pbcd_all.py:121-126
template <int KIND, int DEG, int NA>
__device__ __forceinline__ void sync_one(double (&Av)[NA], double x, double pold, double upd, double pnew,
                                         double lam, double &dy, double &dy2) {
    if (KIND == PK_FM) {
        double dprev = x;
#pragma unroll
        for (int r = 1; r < DEG; r++) {
            const double Aold = Av[r - 1];
            const double dcur = x * (Aold - pold * dprev);
            Av[r - 1] = Aold - upd * dprev;
            dprev = dcur;
        }
        dy = (lam * upd) * dprev;
        dy2 = 0.0;
    } else {
        double Aval = Av[0];
        dy = lam * Aval;
        Aval = Aval / (1.0 + x * pold);
        Aval = Aval * (1.0 + x * pnew);
        Av[0] = Aval;
        dy2 = lam * Aval;
    }
}

// bulk CTAs that share one column (the window has few columns, the GPU many SMs)
__device__ __forceinline__ int bw_parts(int nbulk, int nb) {
    int p = nbulk / (nb > 0 ? nb : 1);
    return p < 1 ? 1 : (p > BPARTS ? BPARTS : p);
}

// ------------------------------------------------------------------------------------ bulk CTAs
template <int KIND, int DEG>
__device__ void bw_bulk(const BWArgs &a, int b, int nbulk) {
    constexpr int NA = (KIND == PK_FM) ? DEG - 1 : 1;
    __shared__ double red[2][BT / 32][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = BT / 32;
    const int k = a.k;
    const bool act = lane < k;
    const size_t strideA = (size_t)NA * k;
    const double lam = act ? a.lams[lane] : 0.0;
    for (int w = 0; w < a.nwin + a.H; w++) {
        if (w < a.nwin) {
            if (tid == 0) {
                if (w - 1 - a.H >= 0) wait_ge_b(a.wb_cnt + (w - 1 - a.H), nbulk, true);
                wait_ge_b(a.eng_done, w - a.H, true);
            }
            __syncthreads();
            const int t0 = w * a.B, nb = min(a.B, a.d - t0);
            const int parts = bw_parts(nbulk, nb);
            for (int task = b; task < nb * parts; task += nbulk) {
                const int tl = task / parts, part = task % parts;
                const int t = t0 + tl, j = a.idx_feat[t];
                const double pold = act ? a.P[(size_t)j * k + lane] : 0.0;
                const int cs = a.indptr[j], ce = a.indptr[j + 1];
                const int ps = cs + (int)((long long)(ce - cs) * part / parts);
                const int pe = cs + (int)((long long)(ce - cs) * (part + 1) / parts);
                double g = 0.0, h = 0.0;
                // the warp's nonzeros: e = ps + warp + W*q; lane q prefetches (cflag, x) of its q-th one,
                // then 4 record gathers are kept in flight
                const int cnt = pe - ps;
                const int nq = cnt > warp ? (cnt - warp + W - 1) / W : 0;
                for (int qb = 0; qb < nq; qb += 32) {
                    int fl = -1;
                    double xl = 0.0;
                    if (qb + lane < nq) { const int e = ps + warp + W * (qb + lane); fl = a.cflag[e]; xl = a.data[e]; }
                    const int nloc = min(32, nq - qb);
                    for (int q0 = 0; q0 < nloc; q0 += 4) {
                        int fi[4];
                        double2 yy[4];
                        double Av[4][NA], xv[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            fi[u] = __shfl_sync(0xffffffffu, fl, (q0 + u) & 31);
                            xv[u] = sp_shfl(xl, (q0 + u) & 31);
                            if (q0 + u >= nloc) fi[u] = -1;
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (fi[u] >= 0) {
                                yy[u] = __ldcg(reinterpret_cast<const double2 *>(a.yrec + (size_t)fi[u] * 2));
#pragma unroll
                                for (int r = 0; r < NA; r++)
                                    Av[u][r] = act ? __ldcg(a.A + (size_t)fi[u] * strideA + (size_t)r * k + lane) : 0.0;
                            }
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (fi[u] >= 0) {
                                const double dl = sp_dloss_rt(a.loss, yy[u].x, yy[u].y);
                                const double last = dA_last<KIND, DEG, NA>(Av[u], xv[u], pold);
                                if (act) { g += dl * last; h += last * last; }     // pbcd.py:65-67
                            }
                    }
                }
                red[0][warp][lane] = g; red[1][warp][lane] = h;
                __syncthreads();
                if (warp == 0) {
                    double sg = 0.0, sh = 0.0;
#pragma unroll
                    for (int q = 0; q < W; q++) { sg += red[0][q][lane]; sh += red[1][q][lane]; }
                    double *dst = a.base + ((((size_t)(w % BRING) * BWIN + tl) * BPARTS + part) * 2) * 32;
                    __stcg(dst + lane, sg);
                    __stcg(dst + 32 + lane, sh);
                }
                __syncthreads();
            }
            if (tid == 0) { __threadfence(); red_release_add_b(a.base_cnt + w, 1); }
        }
        const int wv = w - a.H;
        if (wv >= 0) {
            if (tid == 0) wait_ge_b(a.eng_done, wv + 1, true);
            __syncthreads();
            const int t0 = wv * a.B, nb = min(a.B, a.d - t0);
            const int parts = bw_parts(nbulk, nb);
            for (int task = b; task < nb * parts; task += nbulk) {
                const int tl = task / parts, part = task % parts;
                const int t = t0 + tl, j = a.idx_feat[t];
                const double upd = act ? __ldcg(a.res + ((size_t)t * 2) * k + lane) : 0.0;
                const double pnew = act ? __ldcg(a.res + ((size_t)t * 2 + 1) * k + lane) : 0.0;
                const bool moved = __any_sync(0xffffffffu, upd != 0.0);
                if (KIND != PK_ALL && !moved) continue;
                const double pold = act ? a.P[(size_t)j * k + lane] : 0.0;
                const int cs = a.indptr[j], ce = a.indptr[j + 1];
                const int ps = cs + (int)((long long)(ce - cs) * part / parts);
                const int pe = cs + (int)((long long)(ce - cs) * (part + 1) / parts);
                const int cnt = pe - ps;
                const int nq = cnt > warp ? (cnt - warp + W - 1) / W : 0;
                for (int qb = 0; qb < nq; qb += 32) {
                    int fl = -1;
                    double xl = 0.0;
                    if (qb + lane < nq) { const int e = ps + warp + W * (qb + lane); fl = a.cflag[e]; xl = a.data[e]; }
                    const int nloc = min(32, nq - qb);
                    for (int q0 = 0; q0 < nloc; q0 += 4) {
                        int fi[4];
                        double Av[4][NA], xv[4], ypv[4], dy[4], dy2[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            fi[u] = __shfl_sync(0xffffffffu, fl, (q0 + u) & 31);
                            xv[u] = sp_shfl(xl, (q0 + u) & 31);
                            if (q0 + u >= nloc) fi[u] = -1;
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            dy[u] = 0.0; dy2[u] = 0.0; ypv[u] = 0.0;
                            if (fi[u] >= 0) {
                                ypv[u] = __ldcg(a.yrec + (size_t)fi[u] * 2);
#pragma unroll
                                for (int r = 0; r < NA; r++)
                                    Av[u][r] = act ? __ldcg(a.A + (size_t)fi[u] * strideA + (size_t)r * k + lane) : 0.0;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (fi[u] >= 0) {
                                sync_one<KIND, DEG, NA>(Av[u], xv[u], pold, upd, pnew, lam, dy[u], dy2[u]);
                                if (act) {
#pragma unroll
                                    for (int r = 0; r < NA; r++)
                                        __stcg(a.A + (size_t)fi[u] * strideA + (size_t)r * k + lane, Av[u][r]);
                                } else { dy[u] = 0.0; dy2[u] = 0.0; }
                            }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            dy[u] = sp_warp_allsum(dy[u]);
                            if (KIND == PK_ALL) dy2[u] = sp_warp_allsum(dy2[u]);
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (fi[u] >= 0 && lane == 0) {
                                double yp = ypv[u] - dy[u];
                                if (KIND == PK_ALL) yp = yp + dy2[u];
                                __stcg(a.yrec + (size_t)fi[u] * 2, yp);
                            }
                    }
                }
            }
            __syncthreads();
            if (tid == 0) { __threadfence(); red_release_add_b(a.wb_cnt + wv, 1); }
        }
    }
}

// ------------------------------------------------------------------------------------ engine CTA
template <int KIND, int DEG>
__device__ void bw_engine(const BWArgs &a, unsigned char *smem_raw, int nbulk) {
    constexpr int NA = (KIND == PK_FM) ? DEG - 1 : 1;
    constexpr int NC = (KIND == PK_FM) ? DEG + 1 : 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k, d = a.d, B = a.B, slotsz = a.slotsz, reg = a.reg, loss = a.loss;
    const bool act = lane < k;
    const size_t strideA = (size_t)NA * k;
    const double mu = sp_mu_rt(loss);
    const double beta = a.beta, gamma = a.gamma, eta = a.eta;
    const double lam = act ? a.lams[lane] : 0.0;
    const bool use_norms = (reg == SP_REG_SQL21 || reg == SP_REG_OMEGACS);

    double *recs = reinterpret_cast<double *>(smem_raw);                       // [slot_cap][slotsz]
    double *ent_x = recs + (size_t)a.slot_cap * slotsz;                        // [BENT*slot_cap]
    int *ent_sd = reinterpret_cast<int *>(ent_x + BENT * (size_t)a.slot_cap);  // [BENT*slot_cap]
    unsigned char *aux = smem_raw + BW_REC_BYTES;
    double *prev = reinterpret_cast<double *>(aux);                            // [BWIN][32] worker -> chain row
    double *rres = prev + BWIN * 32;                                           // [BWIN][32] chain -> worker row
    double *pold_s = rres + BWIN * 32;                                         // [BWIN][32]
    BCell *cell_l2 = reinterpret_cast<BCell *>(pold_s + BWIN * 32);            // [BWIN] {norm of the row, tag}
    BCell *cell_st = cell_l2 + BWIN;                                           // [BWIN] {strength, tag}
    double *norm_s = reinterpret_cast<double *>(cell_st + BWIN);               // [BWIN]
    unsigned long long *mb_res = reinterpret_cast<unsigned long long *>(norm_s + BWIN);
    unsigned long long *mb_wb = mb_res + BWIN;
    int *hp_s = reinterpret_cast<int *>(mb_wb + BWIN);                         // [BWIN+1]
    int *wbflag = hp_s + BWIN + 1;                                             // [BWIN]

    __shared__ double chain_state[2 + SP_MAXDEG + 1];
    for (int i = tid; i < BWIN; i += BT) {
        cell_l2[i].tag = -1; cell_st[i].tag = -1;
        wbflag[i] = 0;
        mbar_init(smem_u32(&mb_res[i]), 1);
        mbar_init(smem_u32(&mb_wb[i]), 1);
    }
    if (tid == 0) {
        chain_state[0] = *a.viol;
#pragma unroll
        for (int t = 0; t < NC; t++) chain_state[1 + t] = a.regstate[t];
    }
    __syncthreads();

    for (int w = 0; w < a.nwin; w++) {
        const int t0 = w * B, nb = min(B, d - t0);
        const int wtag = w + 1;
        const uint32_t wpar = (uint32_t)(w & 1);
        // ---- stage, part 1: static plan data
        const int h0 = a.ht_ptr[t0];
        const int ns = a.n_slots[w];
        const int32_t *srow = a.slot_row + (size_t)w * a.slot_cap;
        for (int q = tid; q < nb * 32; q += BT) {
            const int tl = q >> 5, s = q & 31;
            const int j = a.idx_feat[t0 + tl];
            pold_s[q] = s < k ? a.P[(size_t)j * k + s] : 0.0;
        }
        for (int tl = tid; tl < nb; tl += BT) {
            const int t = t0 + tl, j = a.idx_feat[t];
            norm_s[tl] = use_norms ? a.norms[j] : 0.0;
            hp_s[tl] = a.ht_ptr[t] - h0;
            if (tl == nb - 1) hp_s[nb] = a.ht_ptr[t + 1] - h0;
        }
        {
            const int nh = a.ht_ptr[t0 + nb] - h0;
            for (int e = tid; e < nh; e += BT) { ent_x[e] = a.h_x[h0 + e]; ent_sd[e] = a.h_sd[h0 + e]; }
        }
        if (tid == 0) {
            wait_ge_b(a.base_cnt + w, nbulk, false);
            if (w - 1 - a.H >= 0) wait_ge_b(a.wb_cnt + (w - 1 - a.H), nbulk, false);
        }
        __syncthreads();
        // ---- stage, part 2: hot sample records (warp per slot)
        {
            constexpr int SU = 4;                              // slots gathered at once per warp
            constexpr int NW = BT / 32;
            const int nq = (NA * k + 31) / 32;                 // <= NA (k <= 32)
            for (int s0 = warp; s0 < ns; s0 += NW * SU) {
                int rows[SU];
                double av[SU][NA], yv[SU];
#pragma unroll
                for (int u = 0; u < SU; u++) rows[u] = (s0 + u * NW < ns) ? srow[s0 + u * NW] : -1;
#pragma unroll
                for (int u = 0; u < SU; u++)
                    if (rows[u] >= 0) {
#pragma unroll
                        for (int r = 0; r < NA; r++)
                            av[u][r] = (r < nq && r * 32 + lane < NA * k) ? __ldcg(a.A + (size_t)rows[u] * strideA + r * 32 + lane) : 0.0;
                        yv[u] = lane < 2 ? __ldcg(a.yrec + (size_t)rows[u] * 2 + lane) : 0.0;
                    }
#pragma unroll
                for (int u = 0; u < SU; u++)
                    if (rows[u] >= 0) {
                        double *dst = recs + (size_t)(s0 + u * NW) * slotsz;
#pragma unroll
                        for (int r = 0; r < NA; r++)
                            if (r < nq && r * 32 + lane < NA * k) dst[r * 32 + lane] = av[u][r];
                        if (lane < 2) dst[NA * k + lane] = yv[u];
                    }
            }
        }
        __syncthreads();

        if (warp == 0) {
            // =========================================================== row prox chain, in order
            double viol = chain_state[0], cache[NC];
#pragma unroll
            for (int q = 0; q < NC; q++) cache[q] = chain_state[1 + q];
            for (int tl = 0; tl < nb; tl++) {
                const long long t = t0 + tl;
                const int j0 = a.idx_feat[t];
                BCell cl, cs;
                do { cl = bcell_load(&cell_l2[tl]); } while (cl.tag != t);
                do { cs = bcell_load(&cell_st[tl]); } while (cs.tag != t);
                const double l2 = cl.v;
                double strength = cs.v;
                double pnew = prev[tl * 32 + lane];
                const double pold = pold_s[tl * 32 + lane];
                // ---- prox_bcd + update_cache_pbcd (same statements as pbcd.cu's block step)
                if (reg == SP_REG_L1) {
                    pnew = sp_soft_threshold(pnew, strength);
                } else {
                    double dc[NC + 1];
#pragma unroll
                    for (int u = 0; u <= NC; u++) dc[u] = 0.0;
                    double norm_j = norm_s[tl];
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
                        pnew = pnew * sc;
                    } else {
                        pnew = 0.0;
                    }
                    if (use_norms) {
                        double dn = sp_warp_allsum(pnew * pnew);
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
                                if (lane == 0) a.norms[j0] = l2n;
                                __syncwarp();
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
                        if (lane == 0) a.norms[j0] = l2n;
                    }
                }
                const double upd = pold - pnew;
                double l1 = sp_warp_allsum(fabs(upd));               // pbcd.py:146 norm(updates, 1)
                viol += l1;
                rres[tl * 32 + lane] = (KIND == PK_ALL) ? pnew : upd;
                if (act) {
                    __stcg(a.res + ((size_t)t * 2) * k + lane, upd);
                    __stcg(a.res + ((size_t)t * 2 + 1) * k + lane, pnew);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&mb_res[tl]));
            }
            if (lane == 0) {
                chain_state[0] = viol;
#pragma unroll
                for (int q = 0; q < NC; q++) chain_state[1 + q] = cache[q];
            }
        } else {
            // =========================================================== workers (lane = component)
            const int W = BT / 32 - 1, wk = warp - 1;
            for (int tl = wk; tl < nb; tl += W) {
                const long long t = t0 + tl;
                const int hs = hp_s[tl], ne = hp_s[tl + 1] - hs;
                const double pold = pold_s[tl * 32 + lane];
                double bg = 0.0, bh = 0.0;                             // cold sums: the parts in fixed order
                {
                    const int parts = bw_parts(nbulk, nb);
                    const double *src = a.base + (((size_t)(w % BRING) * BWIN + tl) * BPARTS) * 64;
                    for (int p = 0; p < parts; p++) {
                        bg += __ldcg(src + (size_t)p * 64 + lane);
                        bh += __ldcg(src + (size_t)p * 64 + 32 + lane);
                    }
                }
                double g = 0.0, h = 0.0;
                for (int e0 = 0; e0 < ne; e0 += 2) {                  // two hot nonzeros in flight
                    int slot[2];
                    double xv[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        slot[u] = -1; xv[u] = 0.0;
                        if (e0 + u < ne) {
                            const int sd = ent_sd[hs + e0 + u];
                            xv[u] = ent_x[hs + e0 + u];
                            slot[u] = sd & 0xffff;
                            const int dep = ((sd >> 16) & 0x1ff) - 1;
                            if (dep >= 0 && bflag_load(&wbflag[dep]) != wtag) mbar_wait(smem_u32(&mb_wb[dep]), wpar);
                        }
                    }
                    double Av[2][NA], yp[2], yv[2];
#pragma unroll
                    for (int u = 0; u < 2; u++)
                        if (slot[u] >= 0) {
                            const double *rec = recs + (size_t)slot[u] * slotsz;
#pragma unroll
                            for (int r = 0; r < NA; r++) Av[u][r] = act ? rec[r * k + lane] : 0.0;
                            yp[u] = rec[NA * k]; yv[u] = rec[NA * k + 1];
                        }
#pragma unroll
                    for (int u = 0; u < 2; u++)
                        if (slot[u] >= 0) {
                            const double dl = sp_dloss_rt(loss, yp[u], yv[u]);
                            const double last = dA_last<KIND, DEG, NA>(Av[u], xv[u], pold);
                            if (act) { g += dl * last; h += last * last; }
                        }
                }
                g = g + bg;
                h = h + bh;
                double inv = sp_warp_allsum(h);                       // pbcd.py:68-72
                inv = inv * mu;
                inv = inv + beta;
                double gr = g * lam;                                  // pbcd.py:74-78
                gr = gr + beta * pold;
                gr = gr / inv;
                double pre = pold - eta * gr;
                const double strength = eta * gamma / inv;
                double l2 = 0.0;
                if (reg != SP_REG_L1) {
                    if (reg == SP_REG_SQL21) pre = pre / (1.0 + 2.0 * strength);
                    const double dot = sp_warp_allsum(act ? pre * pre : 0.0);
                    l2 = sqrt(dot);
                }
                prev[tl * 32 + lane] = act ? pre : 0.0;
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    bcell_store(&cell_st[tl], strength, t);
                    bcell_store(&cell_l2[tl], l2, t);
                }
                mbar_wait(smem_u32(&mb_res[tl]), wpar);
                const double rv = rres[tl * 32 + lane];
                double upd, pnew;
                if (KIND == PK_ALL) { pnew = rv; upd = pold - pnew; }
                else { upd = rv; pnew = 0.0; }
                const bool moved = __any_sync(0xffffffffu, upd != 0.0);
                if (KIND == PK_ALL || moved) {
                    for (int e0 = 0; e0 < ne; e0 += 2) {              // (a position's samples are distinct)
                        double *rec[2];
                        double Av[2][NA], dy[2], dy2[2];
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            rec[u] = nullptr; dy[u] = 0.0; dy2[u] = 0.0;
                            if (e0 + u < ne) {
                                rec[u] = recs + (size_t)(ent_sd[hs + e0 + u] & 0xffff) * slotsz;
#pragma unroll
                                for (int r = 0; r < NA; r++) Av[u][r] = act ? rec[u][r * k + lane] : 0.0;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 2; u++)
                            if (rec[u] != nullptr) {
                                sync_one<KIND, DEG, NA>(Av[u], ent_x[hs + e0 + u], pold, upd, pnew, lam, dy[u], dy2[u]);
                                if (act) {
#pragma unroll
                                    for (int r = 0; r < NA; r++) rec[u][r * k + lane] = Av[u][r];
                                } else { dy[u] = 0.0; dy2[u] = 0.0; }
                            }
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            dy[u] = sp_warp_allsum(dy[u]);
                            if (KIND == PK_ALL) dy2[u] = sp_warp_allsum(dy2[u]);
                        }
#pragma unroll
                        for (int u = 0; u < 2; u++)
                            if (rec[u] != nullptr && lane == 0) {
                                double yp = rec[u][NA * k] - dy[u];
                                if (KIND == PK_ALL) yp = yp + dy2[u];
                                rec[u][NA * k] = yp;
                            }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    bflag_store(&wbflag[tl], wtag);
                    mbar_arrive(smem_u32(&mb_wb[tl]));
                }
            }
        }
        __syncthreads();
        // ---- flush the hot records, publish the window
        for (int sl = warp; sl < ns; sl += BT / 32) {
            const int i = srow[sl];
            const double *src = recs + (size_t)sl * slotsz;
            for (int q = lane; q < NA * k; q += 32) __stcg(a.A + (size_t)i * strideA + q, src[q]);
            if (lane == 0) __stcg(a.yrec + (size_t)i * 2, src[NA * k]);
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); st_release_b(a.eng_done, w + 1); }
    }
    if (tid == 0) {
        *a.viol = chain_state[0];
#pragma unroll
        for (int t = 0; t < NC; t++) a.regstate[t] = chain_state[1 + t];
    }
}

template <int KIND, int DEG>
__global__ void __launch_bounds__(BT, 1) pbcd_wsweep_kernel(const BWArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbulk = gridDim.x - 1;
    if (blockIdx.x == 0) bw_engine<KIND, DEG>(a, smem_raw, nbulk);
    else bw_bulk<KIND, DEG>(a, blockIdx.x - 1, nbulk);
}

// P[idx_feat[t], :] = new row of position t
__global__ void apply_rows_kernel(int d, int k, const int32_t *__restrict__ idx_feat, const double *__restrict__ res,
                                  double *P) {
    const long long n = (long long)d * k;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(q / k), s = (int)(q % k);
        P[(size_t)idx_feat[t] * k + s] = res[((size_t)t * 2 + 1) * k + s];
    }
}

int g_sm_count_b = 0;

template <int KIND, int DEG>
int launch_bw(BWArgs a, double *P_out, cudaStream_t st) {
    auto kern = pbcd_wsweep_kernel<KIND, DEG>;
    const size_t smem = (size_t)BW_REC_BYTES + BW_AUX_BYTES;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return sp_check_cuda(e, "cudaFuncSetAttribute(pbcd_wsweep_kernel)");
    if (g_sm_count_b == 0) {
        int dev = 0;
        SP_CUDA(cudaGetDevice(&dev));
        SP_CUDA(cudaDeviceGetAttribute(&g_sm_count_b, cudaDevAttrMultiProcessorCount, dev));
    }
    int nbulk = (a.B < a.d ? a.B : a.d) * BPARTS;
    if (nbulk > g_sm_count_b - 1) nbulk = g_sm_count_b - 1;
    if (nbulk < 1) nbulk = 1;
    SP_CUDA(cudaMemsetAsync(a.base_cnt, 0, sizeof(int) * (2 * (size_t)(a.nwin + 2) + 2), st));
    void *params[] = {(void *)&a};
    sp_prof_begin(SP_PROF_SWEEP_PBCD, st);
    cudaError_t le = cudaLaunchCooperativeKernel((void *)kern, dim3(nbulk + 1), dim3(BT), params, smem, st);
    if (le == cudaSuccess) {
        int blocks = (int)(((long long)a.d * a.k + 255) / 256);
        if (blocks > 1184) blocks = 1184;
        apply_rows_kernel<<<blocks, 256, 0, st>>>(a.d, a.k, a.idx_feat, a.res, P_out);
        le = cudaGetLastError();
    }
    sp_prof_end(st);
    return sp_check_cuda(le, "pbcd_wsweep_kernel launch");
}

}  // namespace

// doubles of one hot slot: A[i, :, :] + {y_pred, y}, padded to an even count
int sp_pbcd_slotsz(int degree, int k) {
    const int na = (degree == -1) ? 1 : degree - 1;
    return (na * k + 2 + 1) & ~1;
}

// doubles of the cold-sum ring the plan's `base` buffer must hold
extern "C" size_t sp_pbcd_wplan_base_doubles(void) { return (size_t)BRING * BWIN * BPARTS * 64; }

extern "C" int sp_pbcd_wplan_slot_cap(int degree, int k) {
    if (k < 1 || k > 32) return 0;                                  // the window engine handles k <= 32
    return BW_REC_BYTES / (sp_pbcd_slotsz(degree, k) * 8 + 12 * BENT);
}

int sp_pbcd_wsweep(const sp_dataset *ds, const sp_wplan *wp, const int32_t *idx_feat, double *P_dk, int k,
                   const double *lams, int degree, double beta, double gamma, double eta, int reg, int loss,
                   double *yrec, double *A, double *norms, double *regstate, double *viol, cudaStream_t st) {
    if (!wp->cflag || !wp->ht_ptr || !wp->h_sd || !wp->h_x || !wp->n_slots || !wp->slot_row || !wp->sync ||
        !wp->res || !wp->base) {
        sp_set_error("pbcd window plan: missing buffers");
        return SP_ERR_INVALID;
    }
    if (k > 32 || wp->window < 1 || wp->window > BWIN || wp->horizon < 0 || wp->horizon > 1 || wp->near != 0 ||
        wp->slot_cap > sp_pbcd_wplan_slot_cap(degree, k)) {
        sp_set_error("pbcd window plan: k %d / window %d / horizon %d / near %d / slot_cap %d invalid", k, wp->window,
                     wp->horizon, wp->near, wp->slot_cap);
        return SP_ERR_INVALID;
    }
    BWArgs a = {};
    a.d = ds->n_features; a.k = k; a.B = wp->window; a.H = wp->horizon; a.nwin = wp->n_windows;
    a.slot_cap = wp->slot_cap; a.slotsz = sp_pbcd_slotsz(degree, k); a.reg = reg; a.loss = loss;
    a.indptr = ds->csc_indptr; a.cflag = wp->cflag; a.data = ds->csc_data; a.idx_feat = idx_feat;
    a.ht_ptr = wp->ht_ptr; a.h_sd = wp->h_sd; a.h_x = wp->h_x; a.n_slots = wp->n_slots; a.slot_row = wp->slot_row;
    a.P = P_dk; a.lams = lams; a.beta = beta; a.gamma = gamma; a.eta = eta;
    a.yrec = yrec; a.A = A; a.norms = norms; a.regstate = regstate; a.viol = viol;
    a.res = wp->res; a.base = wp->base;
    a.base_cnt = wp->sync;
    a.wb_cnt = wp->sync + (a.nwin + 2);
    a.eng_done = wp->sync + 2 * (a.nwin + 2);
    switch (degree) {
    case -1: return launch_bw<PK_ALL, 1>(a, P_dk, st);
    case 2: return launch_bw<PK_FM, 2>(a, P_dk, st);
    case 3: return launch_bw<PK_FM, 3>(a, P_dk, st);
    case 4: return launch_bw<PK_FM, 4>(a, P_dk, st);
    case 5: return launch_bw<PK_FM, 5>(a, P_dk, st);
    }
    sp_set_error("pbcd window sweep: degree %d unsupported", degree);
    return SP_ERR_UNSUPPORTED;
}
