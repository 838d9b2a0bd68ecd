// Pipelined window sweep: the same coordinate descent as pcd.cu (reference optimizer/pcd.py:33-137,
// pcd_all.py:21-102, cd_linear.py:8-33), organised so that the only strictly sequential work per
// coordinate is the ~10-flop regularizer chain.
//
// One persistent cooperative launch per sweep (one CTA per SM):
//   * CTA 0 is the ENGINE.  Per window of B positions it holds the records of the window's hot
//     samples in shared-memory slots (wplan.cu).  Warp 0 runs the scalar chain (prox_cd + cache,
//     in coordinate order, state in registers); warps 1.. are workers, worker k owning positions
//     k, k+W, ... : gradient terms of the position's hot nonzeros (waiting, per nonzero, on the
//     write-back flag of the position that last touched its slot), + the cold partial sums, the
//     Newton step and the two divisions, handed to the chain warp through a 16-byte {value, tag}
//     cell; then the write-back of the hot records once the chain warp has published the update.
//   * CTAs 1.. are BULK workers: BASE(w) reduces the cold nonzeros of window w's columns to
//     (g, h) per position; WB(w) applies the published updates to the cold records.  With horizon
//     H = 1 a cold record is untouched for a whole window on either side, so BASE(w) runs while
//     the engine is still in window w-1 and the engine never waits for the bulk CTAs.
//   * windows are handed over through three counters in global memory (release / acquire).
//   * ZERO-UPDATE SPECULATION (ANOVA sweeps, windows in which at most 55 % of the coordinates start
//     nonzero -- the regime the sparsity-inducing regularizers drive the fit into): a coordinate whose
//     update is 0 changes no record and leaves the regularizer state unchanged, so nothing depends on
//     it.  Workers then wait only for the last KNOWN MOVER (coordinate that starts nonzero) touching
//     each of their records (per-slot masks slot_mv) and guard against SURPRISES (a zero coordinate
//     that moves) with a snapshot of the chain warp's progress C (all record-changing updates decided
//     before C are applied: counter nz_done) sent with their sums; the chain warp -- 32 positions per
//     step, committing up to the first one that moves -- accepts a cell iff no surprise happened in
//     [C, t), else the worker redoes the position once the chain is parked on it.  Bulk CTAs compute
//     BASE(w+1) early; after a window without record-changing updates the engine starts the next one
//     without any bulk hand-over.  Committed values are those of the non-speculative path (same
//     operations on the same record values; only the summation order of a column's hot terms differs).
// Per-sample terms are computed exactly as the reference does; only the order in which a column's
// terms are summed differs (as in any parallel reduction).
#include "common.cuh"
#include "cluster.cuh"
#include "pcd_common.cuh"
#include "sparsepoly_b200.h"

namespace {

constexpr int WT = 384;                        // threads per CTA: chain warp + 11 worker warps (<= 168 regs)
constexpr int W_AUX_BYTES = SP_WINDOW_MAX * (3 * 16 + 2 * 8 + 16 + 4 * 4 + 2 * 8) + 64;
constexpr int W_REC_BYTES = 196608;            // shared memory reserved for the hot record slots

struct WArgs {
    int d, B, H, nwin, slot_cap, stride, reg, spec, spec_denom;
    const int32_t *indptr;       // CSC
    const int32_t *cflag;
    const double *data;
    const int32_t *idx_feat;
    const int32_t *ht_ptr, *ht_cls, *h_sd;
    const double *h_x;
    const int32_t *n_slots, *slot_row;
    const double *prow;          // P[s, :] (or w): read-only during the sweep
    const double *cns;
    const double *lam_ptr;
    double ab, gamma, eta;
    double *rec;
    double *regstate, *viol;
    double2 *res, *base;
    int *base_cnt, *wb_cnt, *eng_done;
    int *early_cnt, *nzwin;      // early (speculative) BASE hand-over, record-changing updates per window
};

struct __align__(16) Cell { double v; long long tag; };
#define SP_ENT_FWD 0x20000000       // packed hot nonzero (wplan.cu): slot | (dep+1) << 16 | fwd | late
#define SP_ENT_LATE 0x40000000
#ifndef SP_BACKOFF_NS
#define SP_BACKOFF_NS 0
#endif
#define SP_BACKOFF if (SP_BACKOFF_NS) __nanosleep(SP_BACKOFF_NS);

// Optional cycle accounting of the engine roles (build with -DSP_WPROF; read with sp_wprof_read).
enum { TP_ENG_WAIT = 0, TP_ENG_STAGE, TP_ENG_ROLE, TP_ENG_FLUSH, TP_CH_WAIT, TP_CH_COMP, TP_WK_LOAD, TP_WK_DEP,
       TP_WK_TERMS, TP_WK_RED, TP_WK_RESWAIT, TP_WK_WB, TP_BULK_WAITB, TP_BULK_BASE, TP_BULK_WAITW, TP_BULK_WB,
       TP_N };
__device__ unsigned long long g_wprof[TP_N];
__device__ unsigned long long g_wspec[2];            // speculated positions, rejected speculations
__device__ long long g_wtrace[SP_WINDOW_MAX * 8];   // per-position timestamps of window 100 (SP_WPROF)
#if defined(SP_WPROF) && SP_WPROF >= 2
#define TP_DECL unsigned long long tp_acc[TP_N] = {}; long long tp_t0 = clock64();
#define TP_MARK(id) { const long long tp_t1 = clock64(); tp_acc[id] += (unsigned long long)(tp_t1 - tp_t0); tp_t0 = tp_t1; }
#define TP_FLUSH(cond) if (cond) { for (int i_ = 0; i_ < TP_N; i_++) if (tp_acc[i_]) atomicAdd(&g_wprof[i_], tp_acc[i_]); }
#else
#define TP_DECL
#define TP_MARK(id) {}
#define TP_FLUSH(cond) {}
#endif
#ifndef SP_TRACE_DEG
#define SP_TRACE_DEG DEG            /* trace only the sweeps of this degree (debug builds) */
#endif
#ifdef SP_WPROF
#define TR(w_, tl_, k_) if (DEG == SP_TRACE_DEG && (w_) == 100 && lane == 0) g_wtrace[(tl_) * 8 + (k_)] = clock64();
#define TRD(w_, tl_, k_, v_) { asm volatile("" ::"d"(v_) : "memory"); TR(w_, tl_, k_) }
#else
#define TR(w_, tl_, k_) {}
#define TRD(w_, tl_, k_, v_) {}
#endif

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(int *p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_ge(const int *p, int v) {
    while (ld_acquire(p) < v) {}
}
// bulk CTAs: poll with a back-off (147 CTAs hammering one L2 line slow the engine's own traffic)
__device__ __forceinline__ void wait_ge_sleep(const int *p, int v) {
    while (ld_acquire(p) < v) __nanosleep(64);
}
__device__ __forceinline__ Cell cell_load(const Cell *c) {
    Cell r;
    unsigned long long a, b;
    asm volatile("ld.volatile.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(smem_u32(c)) : "memory");
    r.v = __longlong_as_double((long long)a);
    r.tag = (long long)b;
    return r;
}
__device__ __forceinline__ void cell_store(Cell *c, double v, long long tag) {
    asm volatile("st.volatile.shared.v2.b64 [%0], {%1, %2};" ::"r"(smem_u32(c)),
                 "l"(__double_as_longlong(v)), "l"(tag)
                 : "memory");
}
// same, on precomputed shared-window addresses (saves the cvta sequence in the hot loops)
__device__ __forceinline__ void cell_store_a(uint32_t addr, double v, int tag) {
    asm volatile("st.volatile.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(__double_as_longlong(v)),
                 "l"((long long)tag)
                 : "memory");
}
// cell with the full 64-bit tag: position in the low word, the worker's progress snapshot in the high
__device__ __forceinline__ void cell_load_a64(uint32_t addr, double &v, long long &tag) {
    unsigned long long a, b;
    asm volatile("ld.volatile.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
    v = __longlong_as_double((long long)a);
    tag = (long long)b;
}
__device__ __forceinline__ unsigned long long lds_u64_volatile(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u64_volatile(unsigned long long *p, unsigned long long v) {
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(smem_u32(p)), "l"(v) : "memory");
}
#define SP_CS_EXACT 0x7fff          // "snapshot" of a cell computed without speculation
#define SP_TAG_REDO (-2)            // result cell: speculation rejected, recompute
__device__ __forceinline__ long long cell_tag(int t, int cs) { return (long long)(unsigned)t | ((long long)cs << 32); }
__device__ __forceinline__ double lds_f64_a(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int flag_load(const int *p) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void flag_store(int *p, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

template <int R> __device__ __forceinline__ void load_rec_cg(const double *p, double (&r)[R]) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
    for (int u = 0; u < (R + 1) / 2; u++) {
        const double2 v = __ldcg(q + u);
        r[2 * u] = v.x;
        if (2 * u + 1 < R) r[2 * u + 1] = v.y;
    }
}
// stores the R used doubles; records are padded to an even stride, the pad lane is rewritten
// with whatever was loaded (pad[R] for odd R)
template <int R> __device__ __forceinline__ void store_rec_cg(double *p, const double (&r)[R], double pad) {
    double2 *q = reinterpret_cast<double2 *>(p);
#pragma unroll
    for (int u = 0; u < (R + 1) / 2; u++) {
        double2 v;
        v.x = r[2 * u];
        v.y = (2 * u + 1 < R) ? r[2 * u + 1] : pad;
        __stcg(q + u, v);
    }
}

// ---------------------------------------------------------------------------------- bulk CTAs
// BASE(w): (g, h) sums over the cold nonzeros of every column of window w this CTA owns -> a.base
template <int KIND, int DEG, int LOSS>
__device__ __forceinline__ void bulk_base(const WArgs &a, int w, int b, int nbulk, double (*red)[WT / 32]) {
    constexpr int NA = (KIND == KIND_FM) ? DEG - 1 : (KIND == KIND_ALL ? 1 : 0);
    constexpr int R = 2 + NA;
    constexpr int ND = (KIND == KIND_FM) ? DEG : 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int stride = a.stride;
    const int t0 = w * a.B, nb = min(a.B, a.d - t0);
    for (int tl = b; tl < nb; tl += nbulk) {
        const int t = t0 + tl, j = a.idx_feat[t];
        const double pold = a.prow[j];
        double tg = 0.0, th = 0.0;
        for (int g = a.indptr[j] + tid; g < a.indptr[j + 1]; g += WT) {
            const int fi = a.cflag[g];
            if (fi >= 0) {
                double r[R], dA[ND];
                load_rec_cg<R>(a.rec + (size_t)fi * stride, r);
                nz_terms<KIND, DEG, LOSS, R, ND>(r, a.data[g], pold, dA, tg, th);
            }
        }
        tg = sp_warp_allsum(tg);
        th = sp_warp_allsum(th);
        if (lane == 0) { red[0][warp] = tg; red[1][warp] = th; }
        __syncthreads();
        if (tid == 0) {
            double sg = 0.0, sh = 0.0;
#pragma unroll
            for (int q = 0; q < WT / 32; q++) { sg += red[0][q]; sh += red[1][q]; }
            __stcg(a.base + t, make_double2(sg, sh));
        }
        __syncthreads();
    }
}

// With horizon 0 and speculation enabled the bulk CTAs also compute BASE(w+1) EARLY, while the engine
// is still in window w, on the assumption that window w changes no record (all its updates are 0: the
// sparse regime).  The engine publishes the number of record-changing updates of a window (nzwin)
// with its hand-over; if it is 0 the early sums stand and the engine starts window w+1 without
// waiting for any bulk work, otherwise WB(w) and an exact BASE(w+1) follow as usual.
template <int KIND, int DEG, int LOSS>
__device__ void bulk_role(const WArgs &a, int b, int nbulk) {
    constexpr int NA = (KIND == KIND_FM) ? DEG - 1 : (KIND == KIND_ALL ? 1 : 0);
    constexpr int R = 2 + NA;
    constexpr int ND = (KIND == KIND_FM) ? DEG : 1;
    __shared__ double red[2][WT / 32];
    __shared__ int nz_s;
    const int tid = threadIdx.x;
    const int stride = a.stride;
    const double lam = (KIND == KIND_LINEAR) ? 1.0 : *a.lam_ptr;
    const bool early = a.H == 0 && a.spec;
    bool early_valid = false;                 // a.base of window w already holds valid early sums
    TP_DECL
    for (int w = 0; w < a.nwin + a.H; w++) {
        if (w < a.nwin) {
            if (!early_valid) {
                // ---- BASE(w): cold records of window w were last written in window <= w-H-1
                if (tid == 0) {
                    if (w - 1 - a.H >= 0) wait_ge_sleep(a.wb_cnt + (w - 1 - a.H), nbulk);
                    wait_ge_sleep(a.eng_done, w - a.H);
                }
                __syncthreads();
                TP_MARK(TP_BULK_WAITB)
                bulk_base<KIND, DEG, LOSS>(a, w, b, nbulk, red);
                if (tid == 0) { __threadfence(); red_release_add(a.base_cnt + w, 1); }
                TP_MARK(TP_BULK_BASE)
            }
            if (early && w + 1 < a.nwin) {
                // ---- early BASE(w+1): everything up to window w-1 is applied (waited for above, or
                //      nothing was written since)
                bulk_base<KIND, DEG, LOSS>(a, w + 1, b, nbulk, red);
                if (tid == 0) { __threadfence(); red_release_add(a.early_cnt + w + 1, 1); }
                TP_MARK(TP_BULK_BASE)
            }
        }
        const int wv = w - a.H;
        if (wv >= 0) {
            // ---- WB(wv): apply the published updates to the cold records of window wv
            if (tid == 0) {
                wait_ge_sleep(a.eng_done, wv + 1);
                nz_s = early ? ld_acquire(a.nzwin + wv) : 1;
            }
            __syncthreads();
            TP_MARK(TP_BULK_WAITW)
            early_valid = early && nz_s == 0;
            const int t0 = wv * a.B, nb = min(a.B, a.d - t0);
            for (int tl = b; tl < nb && !early_valid; tl += nbulk) {
                const int t = t0 + tl, j = a.idx_feat[t];
                const double2 rs = __ldcg(a.res + t);
                const double upd = rs.x, pnew = rs.y;
                if (KIND != KIND_ALL && upd == 0.0) continue;
                const double pold = a.prow[j];
                for (int g = a.indptr[j] + tid; g < a.indptr[j + 1]; g += WT) {
                    const int fi = a.cflag[g];
                    if (fi >= 0) {
                        double r[R], dA[ND];
                        double *p = a.rec + (size_t)fi * stride;
                        load_rec_cg<R>(p, r);
                        const double pad = (R & 1) ? __ldcg(p + R) : 0.0;
                        const double x = a.data[g];
                        nz_dA<KIND, DEG, R, ND>(r, x, pold, dA);
                        nz_update<KIND, DEG, R, ND>(r, dA, x, lam, upd, pold, pnew);
                        store_rec_cg<R>(p, r, pad);
                    }
                }
            }
            __syncthreads();
            if (tid == 0) { __threadfence(); red_release_add(a.wb_cnt + wv, 1); }
            TP_MARK(TP_BULK_WB)
        }
    }
    TP_FLUSH(tid == 0 && b == 0)
}

// ---------------------------------------------------------------------------------- engine CTA
template <int KIND, int DEG, int LOSS>
__device__ void engine_role(const WArgs &a, unsigned char *smem_raw, int nbulk) {
    constexpr int NA = (KIND == KIND_FM) ? DEG - 1 : (KIND == KIND_ALL ? 1 : 0);
    constexpr int R = 2 + NA;
    constexpr int NCH = (R + 1) / 2;
    constexpr int ND = (KIND == KIND_FM) ? DEG : 1;
    constexpr int NC = (KIND == KIND_FM) ? DEG : 1;
    constexpr int BM = SP_WINDOW_MAX;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int stride = a.stride, d = a.d, B = a.B;
    const double lam = (KIND == KIND_LINEAR) ? 1.0 : *a.lam_ptr;
    const double mu = sp_mu<LOSS>();
    const double ab = a.ab, gamma = a.gamma, eta = a.eta;

    double *recs = reinterpret_cast<double *>(smem_raw);                       // [slot_cap*stride]
    double *ent_x = recs + (size_t)a.slot_cap * stride;                        // [2*slot_cap] values
    int *ent_sd = reinterpret_cast<int *>(ent_x + 2 * (size_t)a.slot_cap);     // [2*slot_cap] slot | (dep+1)<<16
    unsigned *slot_mv = reinterpret_cast<unsigned *>(ent_sd + 2 * (size_t)a.slot_cap);   // [slot_cap][4] mover masks
    Cell *cellA = reinterpret_cast<Cell *>(smem_raw + W_REC_BYTES);            // [BM] worker -> chain
    Cell *cellB = cellA + BM;                                                  // [BM]
    Cell *rcell = cellB + BM;                                                  // [BM] chain -> worker
    double2 *base_s = reinterpret_cast<double2 *>(rcell + BM);                 // [BM]
    double *pold_s = reinterpret_cast<double *>(base_s + BM);                  // [BM]
    double *cn_s = pold_s + BM;                                                // [BM]
    unsigned long long *mb_res = reinterpret_cast<unsigned long long *>(cn_s + BM);   // [BM] result published
    unsigned long long *mb_wb = mb_res + BM;                                   // [BM] write-back done
    int *hp_s = reinterpret_cast<int *>(mb_wb + BM);                           // [BM+1]
    int *wbflag = hp_s + BM + 1;                                               // [BM]
    int *cls_s = wbflag + BM;                                                  // [BM] chain-warp nonzero counts

    for (int i = tid; i < BM; i += WT) {
        cellA[i].tag = -1; cellB[i].tag = -1; rcell[i].tag = -1;
        wbflag[i] = 0;
        mbar_init(smem_u32(&mb_res[i]), 1);
        mbar_init(smem_u32(&mb_wb[i]), 1);
    }
    // chain-warp state (viol, regularizer cache): lives in shared memory between windows so that it
    // only occupies registers inside the chain loop
    __shared__ double chain_state[1 + SP_MAXDEG];
    __shared__ unsigned long long prog_s;      // chain progress in this window: decided positions | nonzero updates << 32
    __shared__ int nzdone_s;                   // nonzero updates of this window whose write-back is complete
    __shared__ int specoff_s;                  // too many rejections: the rest of the window runs exactly
    __shared__ int winnz_s;                    // record-changing updates of the window just finished
    __shared__ unsigned long long spec_cnt_s[2];
    if (tid == 0) {
        spec_cnt_s[0] = 0; spec_cnt_s[1] = 0;
        winnz_s = 1;
        chain_state[0] = *a.viol;
#pragma unroll
        for (int t = 0; t < NC; t++) chain_state[1 + t] = (KIND == KIND_LINEAR) ? 0.0 : a.regstate[t];
    }
    __syncthreads();
    TP_DECL

    for (int w = 0; w < a.nwin; w++) {
        const int t0 = w * B, nb = min(B, d - t0);
        const int wtag = w + 1;
        // ---- stage, part 1 (static plan data: issued before waiting for the bulk CTAs): per-position
        //      scalars, the hot nonzeros, the sample index of every slot
        const int h0 = a.ht_ptr[t0];
        const int ns = a.n_slots[w];
        const int32_t *srow = a.slot_row + (size_t)w * a.slot_cap;
        constexpr int SU = 4;                                // independent gathers in flight per thread
        int rows0[SU];
#pragma unroll
        for (int u = 0; u < SU; u++) {
            const int sl = u * WT + tid;
            rows0[u] = sl < ns ? srow[sl] : -1;
        }
        int my_nz = 0;
        for (int tl = tid; tl < nb; tl += WT) {
            const int t = t0 + tl, j = a.idx_feat[t];
            const double pv = a.prow[j];
            pold_s[tl] = pv;
            my_nz |= pv != 0.0;
            cn_s[tl] = (KIND == KIND_LINEAR) ? a.cns[j] : 0.0;
            hp_s[tl] = a.ht_ptr[t] - h0;
            cls_s[tl] = a.ht_cls[t];
            if (tl == nb - 1) hp_s[nb] = a.ht_ptr[t + 1] - h0;
        }
        {
            const int nh = a.ht_ptr[t0 + nb] - h0;
            constexpr int EU = 4;
            for (int e0 = 0; e0 < nh; e0 += WT * EU) {
                double xv[EU];
                int sv[EU];
#pragma unroll
                for (int u = 0; u < EU; u++) {
                    const int e = e0 + u * WT + tid;
                    if (e < nh) { xv[u] = a.h_x[h0 + e]; sv[u] = a.h_sd[h0 + e]; }
                }
#pragma unroll
                for (int u = 0; u < EU; u++) {
                    const int e = e0 + u * WT + tid;
                    if (e < nh) { ent_x[e] = xv[u]; ent_sd[e] = sv[u]; }
                }
            }
        }
        if (tid == 0) {
            prog_s = 0ull; nzdone_s = 0; specoff_s = 0;
            if (a.H == 0 && a.spec && w > 0 && winnz_s == 0) {
                // the previous window changed no record: the early sums of the bulk CTAs stand
                wait_ge(a.early_cnt + w, nbulk);
            } else {
                wait_ge(a.base_cnt + w, nbulk);
                if (w - 1 - a.H >= 0) wait_ge(a.wb_cnt + (w - 1 - a.H), nbulk);
            }
        }
        // speculate when (almost) every coordinate of the window starts at zero: under l1 / squaredl12 /
        // omegati such coordinates nearly always stay there (nb <= WT: one position per thread)
        const int nz_old = __syncthreads_count(my_nz);
        // measured at C2 (windows of 96, profiles/r01b_bench_pcd_launches_summary.txt): speculative windows win
        // below ~55 % known movers (13 % at 45 %, 2.7x in all-zero sweeps) and lose above (60-66 %: 87-96 ms per
        // sweep against 82 ms): there the plain path with its chain-warp shortcut for adjacent positions is better
        const bool win_spec = KIND == KIND_FM && a.spec && nb <= 128 &&
                              (a.spec_denom ? nz_old * a.spec_denom <= nb : nz_old * 20 <= nb * 11);
        if (win_spec) {
            // slot_mv[slot] = bit mask of the window positions that touch the slot AND start nonzero (known
            // movers: they will change the record).  A speculating worker waits for the write-back of the
            // last such position before its own -- the only true dependency it can know in advance.
            for (int i = tid; i < ns * 4; i += WT) slot_mv[i] = 0u;
            __syncthreads();
            for (int tl = warp; tl < nb; tl += WT / 32) {
                if (pold_s[tl] == 0.0) continue;
                const int hs = hp_s[tl], ne = hp_s[tl + 1] - hs;
                for (int e = lane; e < ne; e += 32)
                    atomicOr(&slot_mv[(ent_sd[hs + e] & 0xffff) * 4 + (tl >> 5)], 1u << (tl & 31));
            }
        }
        if (tid == 32) TP_MARK(TP_ENG_WAIT)
        // ---- stage, part 2 (data the bulk CTAs produce): cold partial sums, hot sample records
        for (int tl = tid; tl < nb; tl += WT) base_s[tl] = __ldcg(a.base + t0 + tl);
        for (int s0 = 0; s0 < ns; s0 += WT * SU) {
            int rows[SU];
            double2 v[SU][NCH];
#pragma unroll
            for (int u = 0; u < SU; u++) {
                const int sl = s0 + u * WT + tid;
                rows[u] = (s0 == 0) ? rows0[u] : (sl < ns ? srow[sl] : -1);
            }
#pragma unroll
            for (int u = 0; u < SU; u++)
                if (rows[u] >= 0) {
                    const double2 *src = reinterpret_cast<const double2 *>(a.rec + (size_t)rows[u] * stride);
#pragma unroll
                    for (int h = 0; h < NCH; h++) v[u][h] = __ldcg(src + h);
                }
#pragma unroll
            for (int u = 0; u < SU; u++)
                if (rows[u] >= 0) {
                    double2 *dst = reinterpret_cast<double2 *>(recs + (size_t)(s0 + u * WT + tid) * stride);
#pragma unroll
                    for (int h = 0; h < NCH; h++) dst[h] = v[u][h];
                }
        }
        __syncthreads();
        if (tid == 32) TP_MARK(TP_ENG_STAGE)
        if (tid != 32) TP_MARK(TP_ENG_WAIT)

        if (warp == 0 && KIND == KIND_FM && win_spec) {
            // =========================================================== scalar chain, 32 positions per step
            // A coordinate that stays at zero leaves the regularizer state bit-for-bit unchanged, so in a
            // sparse window the chain steps of consecutive positions are independent: lane l evaluates
            // position tl+l against the current state; everything up to and including the first position
            // that moves (update != 0 or state changed) is committed, the rest is re-evaluated next round.
            double viol = chain_state[0], cache[NC];
#pragma unroll
            for (int q = 0; q < NC; q++) cache[q] = chain_state[1 + q];
            const int reg = a.reg;
            // last_nz: last SURPRISE (a coordinate that started at zero and moved); known movers are waited
            // for by the workers (slot_mv), so only surprises invalidate a speculative evaluation
            int last_nz = -1, nz_issued = 0, n_rej = 0, redo_tl = -1;
            unsigned long long n_spec = 0;
            int tl = 0;
            while (tl < nb) {
                const int my = tl + lane, t = t0 + my;
                const bool inw = my < nb;
                double va = 0.0, vb = 0.0;
                long long ta = -1, tb = -1;
                if (inw) {
                    cell_load_a64(smem_u32(&cellA[my]), va, ta);      // (the worker stores B before A)
                    cell_load_a64(smem_u32(&cellB[my]), vb, tb);
                }
                const int cs = (int)(ta >> 32);
                bool ready = inw && (int)ta == t && (int)tb == t && (int)(tb >> 32) == cs;
                if (my == redo_tl) ready = ready && cs == SP_CS_EXACT;
                const unsigned rmask = __ballot_sync(0xffffffffu, ready);
                int n = (rmask == 0xffffffffu) ? 32 : __ffs(~rmask) - 1;     // ready run starting at lane 0
                if (n == 0) continue;
                const unsigned rejmask = __ballot_sync(0xffffffffu, lane < n && last_nz >= cs);
                if (rejmask) {
                    const int r = __ffs(rejmask) - 1;
                    if (r == 0) {
                        // position tl was evaluated against records a later-decided update changed: redo
                        if (lane == 0) {
                            cell_store(&rcell[tl], 0.0, SP_TAG_REDO);
                            mbar_arrive(smem_u32(&mb_res[tl]));
                            if (++n_rej == 4) flag_store(&specoff_s, 1);
                        }
                        n_spec += 1;
                        redo_tl = tl;
                        continue;
                    }
                    n = r;
                }
                const double pold = inw ? pold_s[my] : 0.0;
                double mc[NC];
#pragma unroll
                for (int q = 0; q < NC; q++) mc[q] = cache[q];
                const double pnew = prox_chain<KIND, DEG, NC>(reg, va, vb, pold, mc);
                const double upd = pold - pnew;                      // pcd.py:121
                bool same = true;
#pragma unroll
                for (int q = 0; q < NC; q++) same = same && (__double_as_longlong(mc[q]) == __double_as_longlong(cache[q]));
                const unsigned mmask = __ballot_sync(0xffffffffu, lane < n && !(upd == 0.0 && same));
                const int f = mmask ? __ffs(mmask) - 1 : -1;
                const int c = mmask ? f + 1 : n;
#ifdef SP_WPROF
                if (DEG == SP_TRACE_DEG && w == 100 && lane < c) {
                    g_wtrace[my * 8 + 4] = clock64();
                    g_wtrace[my * 8 + 3] = (long long)c;          // (size of the committed run, not a time)
                }
#endif
                if (lane < c) {
                    cell_store(&rcell[my], upd, t);
                    if (my != redo_tl) mbar_arrive(smem_u32(&mb_res[my]));
                    __stcg(a.res + t, make_double2(upd, pnew));
                }
                n_spec += (unsigned long long)__popc(__ballot_sync(0xffffffffu, lane < c && cs != SP_CS_EXACT && my != redo_tl));
                if (mmask) {
#pragma unroll
                    for (int q = 0; q < NC; q++) cache[q] = sp_shfl(mc[q], f);
                    const double uf = sp_shfl(upd, f);
                    const double pf = sp_shfl(pold, f);
                    if (uf != 0.0) {
                        nz_issued++;
                        if (pf == 0.0) last_nz = tl + f;
                    }
                    viol += fabs(uf);                                // (the committed zeros add +0.0)
                }
                __syncwarp();
                if (lane == 0)
                    sts_u64_volatile(&prog_s, (unsigned long long)(unsigned)(tl + c) | ((unsigned long long)(unsigned)nz_issued << 32));
                tl += c;
            }
            if (lane == 0) {
                chain_state[0] = viol;
#pragma unroll
                for (int q = 0; q < NC; q++) chain_state[1 + q] = cache[q];
                spec_cnt_s[0] += n_spec;
                spec_cnt_s[1] += (unsigned long long)n_rej;
                winnz_s = nz_issued;
            }
        } else if (warp == 0) {
            // =========================================================== scalar chain, in order
            // (windows that do not speculate: every cell was computed after the per-record waits)
            double viol = chain_state[0], cache[NC];
#pragma unroll
            for (int q = 0; q < NC; q++) cache[q] = chain_state[1 + q];
            const int reg = a.reg;
            uint32_t pa = smem_u32(cellA), pb = smem_u32(cellB), pr = smem_u32(rcell), pp = smem_u32(pold_s),
                     pm = smem_u32(mb_res);
            double2 *resg = a.res + t0;
            const bool lane0 = lane == 0;
            double va, vb = 0.0, pold;
            long long ta, tb;
            cell_load_a64(pa, va, ta);
            cell_load_a64(pb, vb, tb);
            pold = lds_f64_a(pp);
            int nz_issued = 0;
            for (int tl = 0; tl < nb; tl++) {
                const int t = t0 + tl;
                // ---- this position's chain nonzeros, one per lane: [late only][late+fwd][fwd only].
                //      Late terms read records the PREVIOUS steps of this warp just wrote back.
                const int cls = cls_s[tl];
                const int nLo = cls & 0xff, nL = nLo + ((cls >> 8) & 0xff), nC = nL + ((cls >> 16) & 0xff);
                double tgl = 0.0, thl = 0.0, cx = 0.0, cr[R], cdA[ND];
                int cslot = -1;
                if (nL > 0) {
                    // (a late record was last written by this warp, so it can be read before the
                    // worker's cell arrives)
                    if (lane < nL) {
                        const int e = hp_s[tl] + lane;
                        cslot = ent_sd[e] & 0xffff;
                        cx = ent_x[e];
                        load_rec<R>(recs + (size_t)cslot * stride, cr);
                        nz_terms<KIND, DEG, LOSS, R, ND>(cr, cx, pold, cdA, tgl, thl);
                    }
                    if (nL > 1) {
                        int P2 = 2;
                        while (P2 < nL) P2 <<= 1;
                        for (int m = P2 >> 1; m > 0; m >>= 1) {
                            tgl += sp_shfl_xor(tgl, m);
                            if (KIND != KIND_LINEAR) thl += sp_shfl_xor(thl, m);
                        }
                    }
                    if (nL > 0) {
                        tgl = sp_shfl(tgl, 0);
                        if (KIND != KIND_LINEAR) thl = sp_shfl(thl, 0);
                    }
                }
                const double cn = (KIND == KIND_LINEAR && nL > 0) ? cn_s[tl] : 0.0;
                while ((int)ta != t) cell_load_a64(pa, va, ta);
                if (KIND != KIND_LINEAR) {
                    while ((int)tb != t) cell_load_a64(pb, vb, tb);
                }
                TR(w, tl, 3)
                // inputs of the next position: in flight while this one is computed (stale tags of an
                // earlier window never match)
                double nva, nvb, npold;
                long long nta, ntb;
                cell_load_a64(pa + 16, nva, nta);
                cell_load_a64(pb + 16, nvb, ntb);
                npold = lds_f64_a(pp + 8);
                if (lane >= nL && lane < nC) {
                    // fwd-only records: their last writer may have been a worker, whose write-back is
                    // ordered before the cell that has just arrived
                    const int e = hp_s[tl] + lane;
                    cslot = ent_sd[e] & 0xffff;
                    cx = ent_x[e];
                    load_rec<R>(recs + (size_t)cslot * stride, cr);
                    nz_dA<KIND, DEG, R, ND>(cr, cx, pold, cdA);
                }
                if (nL > 0) {
                    // the worker sent the sums over the other nonzeros: finish the Newton step here
                    const double tg = va + tgl;
                    if (KIND == KIND_LINEAR) {
                        double u = tg + ab * pold;                   // cd_linear.py:19-22
                        const double inv = mu * cn + ab;
                        va = u / inv;
                    } else {
                        const double th = vb + thl;
                        double inv = th * mu;                        // pcd.py:59-68 / pcd_all.py:34-41
                        inv = inv + ab;
                        double u = tg * lam;
                        u = u + ab * pold;
                        u = u / inv;
                        va = pold - eta * u;
                        vb = eta * gamma / inv;
                    }
                }
                double pnew, upd;
                if (KIND == KIND_LINEAR) {
                    pnew = pold - va;                                // cd_linear.py:24
                    upd = va;
                } else {
                    pnew = prox_chain<KIND, DEG, NC>(reg, va, vb, pold, cache);
                    upd = pold - pnew;                               // pcd.py:121
                }
                if (nC > nLo) {
                    // records the next positions' late terms need: written back right here
                    if (lane >= nLo && lane < nC && (KIND == KIND_ALL || upd != 0.0)) {
                        nz_update<KIND, DEG, R, ND>(cr, cdA, cx, lam, upd, pold, pnew);
                        double *dst = recs + (size_t)cslot * stride;
                        dst[0] = cr[0];
#pragma unroll
                        for (int v = 2; v < R; v++) dst[v] = cr[v];
                    }
                    __syncwarp();
                }
                if (KIND == KIND_ALL || upd != 0.0) nz_issued++;     // record-changing updates of the window
                if (lane0) {
                    cell_store_a(pr, KIND == KIND_ALL ? pnew : upd, t);
                    mbar_arrive(pm);
                    __stcg(resg + tl, make_double2(upd, pnew));
                }
                viol += fabs(upd);
                va = nva; vb = nvb; ta = nta; tb = ntb; pold = npold;
                pa += 16; pb += 16; pr += 16; pp += 8; pm += 8;
                TR(w, tl, 4)
            }
            if (lane0) {
                chain_state[0] = viol;
#pragma unroll
                for (int q = 0; q < NC; q++) chain_state[1 + q] = cache[q];
                winnz_s = nz_issued;
            }
        } else {
            // =========================================================== workers
            constexpr int KR = 2;                               // rounds of 32 hot nonzeros kept in registers
            const int W = WT / 32 - 1, wk = warp - 1;
            const uint32_t wpar = (uint32_t)(w & 1);
            for (int tl = wk; tl < nb; tl += W) {
                const int t = t0 + tl;
                const int hs = hp_s[tl], ne = hp_s[tl + 1] - hs;
                const double pold = pold_s[tl];
                const double2 bs = base_s[tl];
                const double cn = cn_s[tl];
                // In a speculative window the late / fwd classes of the plan are ignored: the workers evaluate
                // and write back every hot nonzero (no waits to shorten, and the chain warp stays minimal).
                const int cls = win_spec ? 0 : cls_s[tl];
                const int nLo = cls & 0xff, nL = nLo + ((cls >> 8) & 0xff);
                const int fwd_mask = win_spec ? 0 : SP_ENT_FWD;
                TR(w, tl, 0)
                int k_slot[KR];                                 // kept for the write-back (-1: none / not ours)
                double k_x[KR], k_r[KR][R], k_dA[KR][ND];
                double upd, pnew;
              for (int attempt = 0;; attempt++) {
                double tg = 0.0, th = 0.0;
#pragma unroll
                for (int u = 0; u < KR; u++) k_slot[u] = -1;
                // speculative (or redone) evaluation: no per-record waits; instead every nonzero update the
                // chain warp had decided when the snapshot was taken must have been applied
                const bool spec = KIND == KIND_FM && (attempt > 0 || (win_spec && flag_load(&specoff_s) == 0));
                int csnap = SP_CS_EXACT;
                if (spec) {
                    unsigned long long pg = 0;
                    if (lane == 0) pg = lds_u64_volatile(&prog_s);
                    pg = __shfl_sync(0xffffffffu, pg, 0);
                    const int nzi = (int)(pg >> 32);
                    while (flag_load(&nzdone_s) < nzi) {}
                    if (attempt == 0) csnap = (int)(pg & 0xffffffffu);
                }
                for (int q0 = 0; q0 < ne; q0 += 32 * KR) {
                    int slot[KR], mydep[KR];
                    double x[KR];
#pragma unroll
                    for (int u = 0; u < KR; u++) {
                        const int e = q0 + u * 32 + lane;
                        slot[u] = -1; mydep[u] = -1; x[u] = 0.0;
                        if (e >= nL && e < ne) {                    // (late nonzeros: the chain warp's)
                            const int sd = ent_sd[hs + e];
                            const int dep = ((sd >> 16) & 0x1ff) - 1;
                            slot[u] = sd & 0xffff;
                            if (sd & fwd_mask) slot[u] |= 0x10000;     // term ours, write-back the chain warp's
                            x[u] = ent_x[hs + e];
                            if (!spec) {
                                if (dep >= 0 && flag_load(&wbflag[dep]) != wtag) mydep[u] = dep;
                            } else if (attempt == 0 && dep >= 0) {
                                // last known mover that touched this slot before this position
                                const uint4 mk = *reinterpret_cast<const uint4 *>(&slot_mv[(sd & 0xffff) * 4]);
                                const unsigned mw[4] = {mk.x, mk.y, mk.z, mk.w};
                                int wi = tl >> 5, lmt = -1;
                                unsigned m = mw[wi] & ((1u << (tl & 31)) - 1u);
                                while (m == 0u && wi > 0) m = mw[--wi];
                                if (m) lmt = (wi << 5) + 31 - __clz(m);
                                if (lmt >= 0 && flag_load(&wbflag[lmt]) != wtag) mydep[u] = lmt;
                            }
                        }
                    }
                    // wait (warp-uniformly: divergent waits would leave the warp fragmented for the
                    // shuffles below) for the write-backs these records depend on
#pragma unroll
                    for (int u = 0; u < KR; u++) {
                        unsigned pending = __ballot_sync(0xffffffffu, mydep[u] >= 0);
                        while (pending) {
                            const int src = __ffs(pending) - 1;
                            const int dpos = __shfl_sync(0xffffffffu, mydep[u], src);
#ifdef SP_SPIN_WAIT
                            while (flag_load(&wbflag[dpos]) != wtag) {}
#else
                            mbar_wait(smem_u32(&mb_wb[dpos]), wpar);
#endif
                            pending &= pending - 1;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < KR; u++) {
                        if (slot[u] >= 0) {
                            double r[R], dA[ND];
                            load_rec<R>(recs + (size_t)(slot[u] & 0xffff) * stride, r);
                            nz_terms<KIND, DEG, LOSS, R, ND>(r, x[u], pold, dA, tg, th);
                            if (q0 == 0 && slot[u] < 0x10000) {
                                k_slot[u] = slot[u]; k_x[u] = x[u];
#pragma unroll
                                for (int v = 0; v < R; v++) k_r[u][v] = r[v];
#pragma unroll
                                for (int v = 0; v < ND; v++) k_dA[u][v] = dA[v];
                            }
                        }
                    }
                }
                TRD(w, tl, 1, tg + th)
                tg = sp_warp_allsum(tg);
                if (KIND != KIND_LINEAR) th = sp_warp_allsum(th);
                TRD(w, tl, 7, tg + th)
                tg = tg + bs.x;
                th = th + bs.y;
                double v0, v1 = 0.0;
                if (nL > 0) {
                    v0 = tg; v1 = th;                                // the chain warp adds its late terms
                } else if (KIND == KIND_LINEAR) {
                    double u = tg + ab * pold;                       // cd_linear.py:19-22
                    const double inv = mu * cn + ab;
                    v0 = u / inv;
                } else {
                    double inv = th * mu;                            // pcd.py:59-68 / pcd_all.py:34-41
                    inv = inv + ab;
                    double u = tg * lam;
                    u = u + ab * pold;
                    u = u / inv;
                    v0 = pold - eta * u;
                    v1 = eta * gamma / inv;
                }
                if (lane == 0) {
                    if (KIND != KIND_LINEAR) cell_store(&cellB[tl], v1, cell_tag(t, csnap));
                    cell_store(&cellA[tl], v0, cell_tag(t, csnap));
                }
                TRD(w, tl, 2, v0 + v1)
                Cell rc;
                if (attempt == 0) {
#ifdef SP_SPIN_WAIT
                    do { rc = cell_load(&rcell[tl]); } while ((int)rc.tag != t && (int)rc.tag != SP_TAG_REDO);
#else
                    mbar_wait(smem_u32(&mb_res[tl]), wpar);
                    rc = cell_load(&rcell[tl]);
#endif
                } else {
                    do { rc = cell_load(&rcell[tl]); } while ((int)rc.tag != t);
                }
                if (KIND == KIND_FM && (int)rc.tag == SP_TAG_REDO) continue;
                TR(w, tl, 5)
                if (KIND == KIND_ALL) { pnew = rc.v; upd = pold - pnew; }
                else { upd = rc.v; pnew = 0.0; }
                break;
              }
                if (KIND == KIND_ALL || upd != 0.0) {
#pragma unroll
                    for (int u = 0; u < KR; u++) {
                        if (k_slot[u] >= 0) {
                            nz_update<KIND, DEG, R, ND>(k_r[u], k_dA[u], k_x[u], lam, upd, pold, pnew);
                            double *dst = recs + (size_t)k_slot[u] * stride;
                            dst[0] = k_r[u][0];
#pragma unroll
                            for (int v = 2; v < R; v++) dst[v] = k_r[u][v];
                        }
                    }
                    // late-only nonzeros (term by the chain warp, write-back ours) and, rarely, nonzeros
                    // beyond the 64 kept in registers: recompute dA from the record
                    for (int e = lane; e < ne; e += 32) {
                        if (e >= nLo && e < 32 * KR) continue;
                        const int sd = ent_sd[hs + e];
                        if (sd & fwd_mask) continue;
                        const int slot = sd & 0xffff;
                        const double x = ent_x[hs + e];
                        double r[R], dA[ND];
                        double *dst = recs + (size_t)slot * stride;
                        load_rec<R>(dst, r);
                        nz_dA<KIND, DEG, R, ND>(r, x, pold, dA);
                        nz_update<KIND, DEG, R, ND>(r, dA, x, lam, upd, pold, pnew);
                        dst[0] = r[0];
#pragma unroll
                        for (int v = 2; v < R; v++) dst[v] = r[v];
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();                      // the write-back above before the flag
                    if (KIND == KIND_ALL || upd != 0.0) atomicAdd_block(&nzdone_s, 1);
                    flag_store(&wbflag[tl], wtag);
                    mbar_arrive(smem_u32(&mb_wb[tl]));          // wakes the warps sleeping on this position
                }
                TR(w, tl, 6)
            }
        }
        __syncthreads();
        if (tid == 32) TP_MARK(TP_ENG_ROLE)
        // ---- flush the hot records (unless no update of this window changed a record), publish the window
        for (int s = tid; s < ns && winnz_s != 0; s += WT) {
            double2 *dst = reinterpret_cast<double2 *>(a.rec + (size_t)srow[s] * stride);
            const double2 *src = reinterpret_cast<const double2 *>(recs + (size_t)s * stride);
#pragma unroll
            for (int h = 0; h < NCH; h++) __stcg(dst + h, src[h]);
        }
        __syncthreads();
        if (tid == 0) {
            a.nzwin[w] = winnz_s;
            __threadfence();
            st_release(a.eng_done, w + 1);
        }
        if (tid == 32) TP_MARK(TP_ENG_FLUSH)
    }
    TP_FLUSH(tid == 0 || tid == 32)
    if (tid == 0) {
        if (spec_cnt_s[0]) atomicAdd(&g_wspec[0], spec_cnt_s[0]);
        if (spec_cnt_s[1]) atomicAdd(&g_wspec[1], spec_cnt_s[1]);
        *a.viol = chain_state[0];
        if (KIND != KIND_LINEAR) {
#pragma unroll
            for (int t = 0; t < NC; t++) a.regstate[t] = chain_state[1 + t];
        }
    }
}

template <int KIND, int DEG, int LOSS>
__global__ void __launch_bounds__(WT, 1) wsweep_kernel(const WArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbulk = gridDim.x - 1;
    if (blockIdx.x == 0) engine_role<KIND, DEG, LOSS>(a, smem_raw, nbulk);
    else bulk_role<KIND, DEG, LOSS>(a, blockIdx.x - 1, nbulk);
}

// prow[idx_feat[t]] = new value of position t (the sweep itself never writes P / w: the bulk
// CTAs read the old values until the last write-back)
__global__ void apply_res_kernel(int d, const int32_t *__restrict__ idx_feat, const double2 *__restrict__ res,
                                 double *prow) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < d; t += gridDim.x * blockDim.x)
        prow[idx_feat[t]] = res[t].y;
}

int g_sm_count = 0;

template <int KIND, int DEG, int LOSS>
int launch_wsweep(WArgs a, double *prow_out, cudaStream_t st) {
    auto kern = wsweep_kernel<KIND, DEG, LOSS>;
    const size_t smem = (size_t)W_REC_BYTES + W_AUX_BYTES;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return sp_check_cuda(e, "cudaFuncSetAttribute(wsweep_kernel)");
    if (g_sm_count == 0) {
        int dev = 0;
        SP_CUDA(cudaGetDevice(&dev));
        SP_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    int nbulk = a.B < a.d ? a.B : a.d;
    if (nbulk > g_sm_count - 1) nbulk = g_sm_count - 1;
    if (nbulk < 1) nbulk = 1;
    SP_CUDA(cudaMemsetAsync(a.base_cnt, 0, sizeof(int) * (4 * (size_t)(a.nwin + 2) + 2), st));
    void *params[] = {(void *)&a};
    sp_prof_begin(SP_PROF_SWEEP_PCD, st);
    cudaError_t le = cudaLaunchCooperativeKernel((void *)kern, dim3(nbulk + 1), dim3(WT), params, smem, st);
    if (le == cudaSuccess) {
        apply_res_kernel<<<(a.d + 255) / 256 > 1184 ? 1184 : (a.d + 255) / 256, 256, 0, st>>>(
            a.d, a.idx_feat, a.res, prow_out);
        le = cudaGetLastError();
    }
    sp_prof_end(st);
    return sp_check_cuda(le, "wsweep_kernel launch");
}

template <int KIND, int DEG>
int wdispatch_loss(int loss, const WArgs &a, double *prow_out, cudaStream_t st) {
    switch (loss) {
    case SP_LOSS_SQUARED: return launch_wsweep<KIND, DEG, SP_LOSS_SQUARED>(a, prow_out, st);
    case SP_LOSS_LOGISTIC: return launch_wsweep<KIND, DEG, SP_LOSS_LOGISTIC>(a, prow_out, st);
    case SP_LOSS_SQHINGE: return launch_wsweep<KIND, DEG, SP_LOSS_SQHINGE>(a, prow_out, st);
    default: sp_set_error("unknown loss id %d", loss); return SP_ERR_INVALID;
    }
}

}  // namespace

// cycle counters of the engine roles since the last call (all zero unless built with -DSP_WPROF)
extern "C" int sp_wprof_read(unsigned long long *out_host /*[16]*/) {
    unsigned long long z[TP_N] = {};
    SP_CUDA(cudaDeviceSynchronize());
    SP_CUDA(cudaMemcpyFromSymbol(out_host, g_wprof, sizeof(z)));
    SP_CUDA(cudaMemcpyToSymbol(g_wprof, z, sizeof(z)));
    return SP_OK;
}

// {positions evaluated speculatively, speculations rejected} since the last call
extern "C" int sp_wspec_read(unsigned long long *out_host /*[2]*/) {
    unsigned long long z[2] = {};
    SP_CUDA(cudaDeviceSynchronize());
    SP_CUDA(cudaMemcpyFromSymbol(out_host, g_wspec, sizeof(z)));
    SP_CUDA(cudaMemcpyToSymbol(g_wspec, z, sizeof(z)));
    return SP_OK;
}

extern "C" int sp_wtrace_read(long long *out_host /*[SP_WINDOW_MAX*8]*/) {
    SP_CUDA(cudaDeviceSynchronize());
    SP_CUDA(cudaMemcpyFromSymbol(out_host, g_wtrace, sizeof(long long) * SP_WINDOW_MAX * 8));
    return SP_OK;
}

extern "C" int sp_wplan_slot_cap(int rec_stride) {
    if (rec_stride < 2) rec_stride = 2;
    // + room for 2 hot nonzeros (12 B each) and a 128-bit mover mask per slot; even, so that the masks
    // (which follow 24*slot_cap bytes of hot nonzeros) stay 16-byte aligned
    return (W_REC_BYTES / (rec_stride * 8 + 40)) & ~1;
}

// degree: 1 = linear (cd_linear), -1 = all-subsets, 2..SP_MAXDEG = ANOVA
int sp_wsweep(const sp_dataset *ds, const sp_wplan *wp, const int32_t *idx_feat, int degree, double *prow,
              const double *cns, const double *lam_ptr, double ab, double gamma, double eta, int reg,
              int loss, double *rec, int rec_stride, double *regstate, double *viol, cudaStream_t st) {
    if (!wp->cflag || !wp->ht_ptr || !wp->ht_cls || !wp->h_sd || !wp->h_x || !wp->n_slots || !wp->slot_row ||
        !wp->sync || !wp->res || !wp->base) {
        sp_set_error("window plan: missing buffers");
        return SP_ERR_INVALID;
    }
    if (wp->window < 1 || wp->window > SP_WINDOW_MAX || wp->horizon < 0 || wp->horizon > 1 ||
        wp->slot_cap > sp_wplan_slot_cap(rec_stride) || wp->slot_cap > 32768 || (wp->slot_cap & 1)) {
        sp_set_error("window plan: window %d / horizon %d / slot_cap %d invalid for record stride %d",
                     wp->window, wp->horizon, wp->slot_cap, rec_stride);
        return SP_ERR_INVALID;
    }
    WArgs a = {};
    a.d = ds->n_features; a.B = wp->window; a.H = wp->horizon; a.nwin = wp->n_windows;
    a.slot_cap = wp->slot_cap; a.stride = rec_stride; a.reg = reg;
    a.spec = (wp->flags & SP_WPLAN_NO_SPECULATION) ? 0 : 1;
    a.spec_denom = (wp->flags >> 8) & 0xff;                   // debug override of the density threshold
    a.indptr = ds->csc_indptr; a.cflag = wp->cflag; a.data = ds->csc_data; a.idx_feat = idx_feat;
    a.ht_ptr = wp->ht_ptr; a.ht_cls = wp->ht_cls; a.h_sd = wp->h_sd; a.h_x = wp->h_x;
    a.n_slots = wp->n_slots; a.slot_row = wp->slot_row;
    a.prow = prow; a.cns = cns; a.lam_ptr = lam_ptr; a.ab = ab; a.gamma = gamma; a.eta = eta;
    a.rec = rec; a.regstate = regstate; a.viol = viol;
    a.res = reinterpret_cast<double2 *>(wp->res);
    a.base = reinterpret_cast<double2 *>(wp->base);
    a.base_cnt = wp->sync;
    a.wb_cnt = wp->sync + (a.nwin + 2);
    a.eng_done = wp->sync + 2 * (a.nwin + 2);
    a.early_cnt = wp->sync + 2 * (a.nwin + 2) + 2;
    a.nzwin = a.early_cnt + (a.nwin + 2);
    switch (degree) {
    case 1: return wdispatch_loss<KIND_LINEAR, 1>(loss, a, prow, st);
    case -1: return wdispatch_loss<KIND_ALL, 1>(loss, a, prow, st);
    case 2: return wdispatch_loss<KIND_FM, 2>(loss, a, prow, st);
    case 3: return wdispatch_loss<KIND_FM, 3>(loss, a, prow, st);
    case 4: return wdispatch_loss<KIND_FM, 4>(loss, a, prow, st);
    case 5: return wdispatch_loss<KIND_FM, 5>(loss, a, prow, st);
    }
    sp_set_error("window sweep: degree %d unsupported", degree);
    return SP_ERR_UNSUPPORTED;
}
