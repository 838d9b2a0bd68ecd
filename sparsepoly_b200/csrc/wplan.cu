// Window plan for the pipelined coordinate sweep (pcd_window.cu).
//
// The coordinate order of pcd / cd_linear (reference optimizer/pcd.py:92-97, cd_linear.py:12) is
// cut into windows of B consecutive positions.  A nonzero (position t, sample i) is HOT when
// sample i has another nonzero in a column visited within `horizon` windows of t's window;
// otherwise it is COLD: nothing else touches that sample's record while the window is in flight.
// Cold nonzeros are reduced / written back in bulk by many CTAs; hot ones go through ONE engine
// CTA that keeps the window's hot sample records in shared memory ("slots") and resolves the
// read-after-write chains between columns exactly, in coordinate order.
//
//   sp_wplan_flag : cflag[e] = row | bit31(hot), hot_count[t] = hot nonzeros of position t
//   sp_wplan_fill : per window: slot table (distinct hot samples), per hot nonzero its slot,
//                   value and `dep` = window-local position that last touched the slot (or -1)
#include "common.cuh"
#include "sparsepoly_b200.h"

namespace {

__global__ void inv_perm_kernel(int d, const int32_t *__restrict__ idx_feat, int32_t *pos) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < d; t += gridDim.x * blockDim.x)
        pos[idx_feat[t]] = t;
}

// one warp per column
__global__ void wflag_kernel(int d, int B, int H, const int32_t *__restrict__ pos,
                             const int32_t *__restrict__ csc_indptr,
                             const int32_t *__restrict__ csc_indices,
                             const int32_t *__restrict__ csr_indptr,
                             const int32_t *__restrict__ csr_indices, int32_t *cflag,
                             int32_t *hot_count) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < d; j += n_warps) {
        const int t = pos[j];
        const int w = t / B;
        int cnt = 0;
        for (int e = csc_indptr[j] + lane; e < csc_indptr[j + 1]; e += 32) {
            const int row = csc_indices[e];
            bool hot = false;
            for (int c = csr_indptr[row]; c < csr_indptr[row + 1]; c++) {
                const int j2 = csr_indices[c];
                if (j2 == j) continue;
                const int w2 = pos[j2] / B;
                const int dw = w2 > w ? w2 - w : w - w2;
                if (dw <= H) { hot = true; break; }
            }
            cflag[e] = (int32_t)((uint32_t)row | (hot ? SP_FLAG_BIT : 0u));
            cnt += hot ? 1 : 0;
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
        if (lane == 0) hot_count[t] = cnt;
    }
}

constexpr int FILL_THREADS = 256;

__device__ __forceinline__ uint32_t hash_row(uint32_t row, int log2hs) {
    return (row * 2654435761u) >> (32 - log2hs);
}

// packed hot nonzero: slot | (dep+1) << 16 | fwd << 29 | late << 30
//   dep  = window-local position that last touched the slot (-1: none)
//   late = dep is at most `near` positions back: the engine's chain warp evaluates this term
//          itself, right after it has produced the update it depends on
//   fwd  = the slot's next toucher is late: the chain warp also writes this record back itself
#define SP_ENT_FWD 0x20000000
#define SP_ENT_LATE 0x40000000

// one CTA per window.  smem: keys[HS] (row or -1), meta[HS] = slot | (last position + 1) << 16,
// lastent[HS] = global index of the slot's previous hot nonzero
__global__ void __launch_bounds__(FILL_THREADS)
wfill_kernel(int d, int B, int slot_cap, int ent_cap, int near, int log2hs, const int32_t *__restrict__ idx_feat,
             const int32_t *__restrict__ csc_indptr, const double *__restrict__ csc_data,
             const int32_t *__restrict__ cflag, const int32_t *__restrict__ ht_ptr, int32_t *tmp_sd,
             double *tmp_x, int32_t *h_sd, double *h_x, int32_t *ht_cls, int32_t *n_slots,
             int32_t *slot_row, int32_t *overflow) {
    extern __shared__ int32_t sm[];
    const int HS = 1 << log2hs;
    int32_t *keys = sm;
    uint32_t *meta = reinterpret_cast<uint32_t *>(sm + HS);
    int32_t *lastent = sm + 2 * HS;
    __shared__ int32_t scan[FILL_THREADS];
    __shared__ int32_t bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = FILL_THREADS / 32;
    const int w = blockIdx.x;
    const int t0 = w * B, nb = min(B, d - t0);
    for (int c = tid; c < HS; c += FILL_THREADS) { keys[c] = -1; meta[c] = 0u; lastent[c] = -1; }
    if (tid == 0) bad = 0;
    __syncthreads();
    // ---- phase 1: compact the hot nonzeros of every position (row order) and collect the rows
    for (int tl = warp; tl < nb; tl += n_warps) {
        const int t = t0 + tl, j = idx_feat[t];
        int out = ht_ptr[t];
        const int s = csc_indptr[j], e = csc_indptr[j + 1];
        for (int base = s; base < e; base += 32) {
            const int g = base + lane;
            const int fi = g < e ? cflag[g] : 0;
            const bool hot = fi < 0;
            const unsigned bal = __ballot_sync(0xffffffffu, hot);
            if (hot) {
                const int o = out + __popc(bal & ((1u << lane) - 1u));
                const int row = fi & SP_ROW_MASK;
                tmp_sd[o] = row;                      // temporary: replaced by the packed entry in phase 2
                tmp_x[o] = csc_data[g];
                uint32_t c = hash_row((uint32_t)row, log2hs);
                bool done = false;
                for (int probe = 0; probe < HS; probe++) {
                    const int32_t prev = atomicCAS(&keys[c], -1, row);
                    if (prev == -1 || prev == row) { done = true; break; }
                    c = (c + 1) & (HS - 1);
                }
                if (!done) bad = 1;
            }
            out += __popc(bal);
        }
    }
    __syncthreads();
    // ---- phase 1b: number the occupied cells (slot ids) in table order
    const int per = HS / FILL_THREADS;
    int mine = 0;
    for (int c = tid * per; c < (tid + 1) * per; c++) mine += keys[c] >= 0;
    scan[tid] = mine;
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int i = 0; i < FILL_THREADS; i++) { const int v = scan[i]; scan[i] = acc; acc += v; }
        n_slots[w] = acc;
        // the engine stages <= slot_cap records and <= ent_cap hot nonzeros per window
        if (acc > slot_cap || bad || ht_ptr[t0 + nb] - ht_ptr[t0] > ent_cap) { *overflow = 1; bad = 1; }
    }
    __syncthreads();
    if (bad) return;                                  // the host retries with a smaller window
    {
        int sl = scan[tid];
        for (int c = tid * per; c < (tid + 1) * per; c++)
            if (keys[c] >= 0) {
                meta[c] = (uint32_t)sl;
                slot_row[(size_t)w * slot_cap + sl] = keys[c];
                sl++;
            }
    }
    __syncthreads();
    // ---- phase 2: positions in order: slot of every hot nonzero, the position that last touched
    //      that slot inside this window, late / fwd marks
    for (int tl = 0; tl < nb; tl++) {
        const int t = t0 + tl;
        for (int o = ht_ptr[t] + tid; o < ht_ptr[t + 1]; o += FILL_THREADS) {
            const int row = tmp_sd[o];
            uint32_t c = hash_row((uint32_t)row, log2hs);
            while (keys[c] != row) c = (c + 1) & (HS - 1);
            const uint32_t mt = meta[c];
            const int dep = (int)(mt >> 16) - 1;
            int sd = (int)(mt & 0xffffu) | ((dep + 1) << 16);
            if (dep >= 0 && tl - dep <= near) {
                sd |= SP_ENT_LATE;
                tmp_sd[lastent[c]] |= SP_ENT_FWD;     // (that nonzero belongs to an earlier position)
            }
            tmp_sd[o] = sd;
            lastent[c] = o;
            meta[c] = (mt & 0xffffu) | ((uint32_t)(tl + 1) << 16);
        }
        __syncthreads();
    }
    // ---- phase 3: order the nonzeros of every position by class (stable): late only, late+fwd,
    //      fwd only, rest -- the chain warp handles the first three, one nonzero per lane
    for (int tl = warp; tl < nb; tl += n_warps) {
        const int t = t0 + tl;
        const int hs = ht_ptr[t], ne = ht_ptr[t + 1] - hs;
        int cnt[4] = {0, 0, 0, 0};
        for (int q = 0; q < ne; q += 32) {
            const int e = q + lane;
            int cls = -1;
            if (e < ne) {
                const int sd = tmp_sd[hs + e];
                const bool late = sd & SP_ENT_LATE, fwd = sd & SP_ENT_FWD;
                cls = late ? (fwd ? 1 : 0) : (fwd ? 2 : 3);
            }
#pragma unroll
            for (int c = 0; c < 4; c++) cnt[c] += __popc(__ballot_sync(0xffffffffu, cls == c));
        }
        int off[4] = {0, cnt[0], cnt[0] + cnt[1], cnt[0] + cnt[1] + cnt[2]};
        for (int q = 0; q < ne; q += 32) {
            const int e = q + lane;
            int cls = -1, sd = 0;
            double x = 0.0;
            if (e < ne) {
                sd = tmp_sd[hs + e]; x = tmp_x[hs + e];
                const bool late = sd & SP_ENT_LATE, fwd = sd & SP_ENT_FWD;
                cls = late ? (fwd ? 1 : 0) : (fwd ? 2 : 3);
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const unsigned bal = __ballot_sync(0xffffffffu, cls == c);
                if (cls == c) {
                    const int o = hs + off[c] + __popc(bal & ((1u << lane) - 1u));
                    h_sd[o] = sd; h_x[o] = x;
                }
                off[c] += __popc(bal);
            }
        }
        if (lane == 0) {
            ht_cls[t] = cnt[0] | (cnt[1] << 8) | (cnt[2] << 16);
            if (cnt[0] + cnt[1] + cnt[2] > 32) { *overflow = 1; }
        }
    }
}

int blocks_for(long long work, int threads) {
    long long b = (work + threads - 1) / threads;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

extern "C" int sp_wplan_flag(const sp_dataset *ds, const int32_t *idx_feat, int window, int horizon,
                             int32_t *pos_scratch, int32_t *cflag, int32_t *hot_count,
                             sp_stream stream) {
    if (!ds || !idx_feat || !pos_scratch || !cflag || !hot_count || !ds->csc_indptr || !ds->csr_indptr ||
        window < 1 || window > SP_WINDOW_MAX || horizon < 0 || horizon > 1) {
        sp_set_error("sp_wplan_flag: invalid argument (window 1..%d, horizon 0..1)", SP_WINDOW_MAX);
        return SP_ERR_INVALID;
    }
    if (ds->n_samples >= (1 << 30)) {
        sp_set_error("sp_wplan_flag: n_samples >= 2^30 is not supported");
        return SP_ERR_UNSUPPORTED;
    }
    const int d = ds->n_features;
    if (d == 0) return SP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    sp_prof_begin(SP_PROF_PLAN, st);
    inv_perm_kernel<<<blocks_for(d, 256), 256, 0, st>>>(d, idx_feat, pos_scratch);
    wflag_kernel<<<blocks_for((long long)d * 32, 256), 256, 0, st>>>(
        d, window, horizon, pos_scratch, ds->csc_indptr, ds->csc_indices, ds->csr_indptr, ds->csr_indices,
        cflag, hot_count);
    sp_prof_end(st);
    SP_LAUNCH_CHECK("wflag_kernel");
    return SP_OK;
}

extern "C" int sp_wplan_fill(const sp_dataset *ds, const int32_t *idx_feat, int window, int slot_cap,
                             int ent_per_slot, int near,
                             const int32_t *cflag, const int32_t *ht_ptr, int32_t *tmp_sd, double *tmp_x,
                             int32_t *h_sd, double *h_x, int32_t *ht_cls, int32_t *n_slots,
                             int32_t *slot_row, int32_t *overflow, sp_stream stream) {
    if (!ds || !idx_feat || !cflag || !ht_ptr || !tmp_sd || !tmp_x || !h_sd || !h_x || !ht_cls || !n_slots ||
        !slot_row || !overflow || window < 1 || window > SP_WINDOW_MAX || slot_cap < 1 || slot_cap > 8192 ||
        near < 0 || near > 16 || ent_per_slot < 1 || ent_per_slot > 8) {
        sp_set_error("sp_wplan_fill: invalid argument");
        return SP_ERR_INVALID;
    }
    const int d = ds->n_features;
    if (d == 0) return SP_OK;
    int log2hs = 9;
    while ((1 << log2hs) < 2 * slot_cap) log2hs++;
    const size_t smem = (size_t)(1 << log2hs) * 12;
    cudaError_t e = cudaFuncSetAttribute(wfill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return sp_check_cuda(e, "cudaFuncSetAttribute(wfill_kernel)");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_windows = (d + window - 1) / window;
    sp_prof_begin(SP_PROF_PLAN, st);
    wfill_kernel<<<n_windows, FILL_THREADS, smem, st>>>(d, window, slot_cap, ent_per_slot * slot_cap, near, log2hs, idx_feat,
                                                        ds->csc_indptr, ds->csc_data, cflag, ht_ptr, tmp_sd,
                                                        tmp_x, h_sd, h_x, ht_cls, n_slots, slot_row, overflow);
    sp_prof_end(st);
    SP_LAUNCH_CHECK("wfill_kernel");
    return SP_OK;
}
