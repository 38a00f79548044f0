// K2 - batched feasibility LP, one LP per group of NW warps, tableau in registers.
//
// Replaces, per candidate active set, MPLP_Program.check_feasibility's LP
//   exists (x,theta):  A x - F theta <= b,  A_t theta <= b_t,  rows of the active set as equalities
// (/root/reference/src/ppopt/mplp_program.py:439-444 -> solver.py:211-246 -> cvxopt_interface.py:153-208),
// the call the serial reference spends 45-50 % of its time in.
// The program's own equalities are already eliminated (host_math.hpp), so the tableau has R0 = mi + q rows and
// nfree + 2 columns [rhs | v | theta | s]; the candidate's active inequality rows are pivoted out first, then
// s = min slack is maximised and the solve stops as soon as s >= -PPG_FEAS_TOL.
#include "common.cuh"
#include "lp_core.cuh"
#include "launch.h"

#include <cstdlib>

namespace ppgpu {

// resident CTAs per SM the register allocator is asked to make room for: the tableau needs 2*RPT*DC registers per
// thread, ~48 more for everything else; more resident LPs hide the barrier / redux latency of a pivot
constexpr int k2_min_blocks(int threads, int rpt, int dc) {
    const int need = 2 * rpt * dc + 48;
    const int mb = 65536 / (threads * need);
    return mb < 1 ? 1 : (mb > 16 ? 16 : mb);
}

template <int NW, int RPT, int DC, int WPC>
__global__ void __launch_bounds__(NW * 32 * WPC, k2_min_blocks(NW * 32 * WPC, RPT, DC))
k2_feas_kernel(DevProgram P, const uint64_t* __restrict__ masks, long long n, uint8_t* __restrict__ status,
               unsigned long long* __restrict__ queue, unsigned long long* __restrict__ counters, int scan_pt) {
    typedef LpCore<NW, RPT, DC> Core;
    constexpr int GT = NW * 32;
    __shared__ typename Core::Shared sh_all[WPC];
    __shared__ long long next_item[WPC];
    __shared__ int need_list[WPC][GT * 16];   // candidates of the scanned block that need an LP (offsets into the block)
    __shared__ int need_cnt[WPC];
    const int grp = (NW == 1) ? (threadIdx.x >> 5) : 0;
    const int tid = (NW == 1) ? (threadIdx.x & 31) : threadIdx.x;
    typename Core::Shared& sh = sh_all[grp];
    unsigned long long n_lp = 0, n_piv = 0, n_work = 0, n_num = 0;
    const int W = P.W;
    // work items: first the candidates K2a handed over with a warm origin (slot list), then the whole level
    long long nwarm = 0;
    if (P.warm_count != nullptr) {
        const unsigned long long c = *P.warm_count;
        nwarm = c < (unsigned long long)P.warm_cap ? (long long)c : P.warm_cap;
    }
    // then the level itself, scanned in blocks of GT * scan_pt status bytes per queue item: after K2a almost nothing is left
    // (one atomic + one status load per CANDIDATE cost more than the LPs of the 2 % that needed one)
    const int blk = GT * scan_pt;
    const long long nblocks = (n + blk - 1) / blk;
    const long long total = nwarm + nblocks;
    auto group_sync = [&]() { if constexpr (NW == 1) __syncwarp(); else __syncthreads(); };
    for (;;) {
        long long item;
        if constexpr (NW == 1) {
            unsigned long long v = 0;
            if (tid == 0) v = atomicAdd(queue, 1ull);
            item = (long long)__shfl_sync(PPG_FULL, v, 0);
        } else {
            if (tid == 0) next_item[0] = (long long)atomicAdd(queue, 1ull);
            __syncthreads();
            item = next_item[0];
            __syncthreads();
        }
        if (item >= total) break;
        const bool warm = item < nwarm;
        const long long base = warm ? 0 : (item - nwarm) * blk;
        int nlist = 1;
        if (!warm) {
            if (tid == 0) need_cnt[grp] = 0;
            group_sync();
            const long long c0 = base + (long long)tid * scan_pt;
            uint8_t sb[16];
            if (scan_pt == 16 && c0 + 16 <= n && (reinterpret_cast<uintptr_t>(status + c0) & 15) == 0) {
                const uint4 v = *reinterpret_cast<const uint4*>(status + c0);
                const unsigned w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 16; ++j) sb[j] = (uint8_t)(w4[j >> 2] >> ((j & 3) * 8));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) sb[j] = (j < scan_pt && c0 + j < n) ? status[c0 + j] : 0;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                // needs an LP: full rank, not certified by K2a, not waiting in the warm list (PPG_ST_PRE)
                if ((sb[j] & PPG_ST_RANK) && !(sb[j] & (PPG_ST_FEAS | PPG_ST_PRE)))
                    need_list[grp][atomicAdd(&need_cnt[grp], 1)] = tid * scan_pt + j;
            }
            group_sync();
            nlist = need_cnt[grp];
        }
        for (int li = 0; li < nlist; ++li) {
        const long long idx = warm ? P.warm_idx[item] : base + need_list[grp][li];
        if (idx < 0) continue;   // slot K2a could not fill (non-finite iterate): that candidate goes the cold way
        const double* wres = warm ? P.warm_resid + (size_t)item * P.R0 : nullptr;
        const uint8_t st = status[idx];
        const uint64_t* mk = masks + idx * W;
        double T[RPT][DC];
        int rflag[RPT];
        const int js = P.nfree + 1;
        static_for<RPT>([&](auto RR) {
            constexpr int rr = decltype(RR)::value;
            const int row = rr * GT + tid;
            if (row < P.R0) {
                const double* src = P.T0 + (size_t)row * P.dc0;
#pragma unroll
                for (int c = 0; c < DC; ++c) T[rr][c] = (c < P.dc0) ? __ldg(src + c) : 0.0;
                if (warm) T[rr][0] = -wres[row];   // h - G z*: the same LP seen from K2a's last iterate
                rflag[rr] = 1;
                if (row < P.mi && mask_test(mk, row)) {
                    rflag[rr] = 2;
                    reg_zero<DC>(T[rr], js);
                }
            } else {
#pragma unroll
                for (int c = 0; c < DC; ++c) T[rr][c] = 0.0;
                rflag[rr] = 0;
            }
        });
        LpOut res = Core::solve(sh, T, rflag, P.R0, js, -PPG_FEAS_TOL, false, tid);
        const bool feas = res.code == PPG_LP_EARLY || res.code == PPG_LP_UNBOUNDED ||
                          (res.code == PPG_LP_OPTIMAL && res.beta >= -PPG_FEAS_TOL);
        if (tid == 0) {
            uint8_t s2 = st;   // a hand-over mark (PPG_ST_PRE) stays: api.cu clears it after this launch
            if (feas) s2 |= PPG_ST_FEAS;
            if (res.code == PPG_LP_ITERLIM) { s2 |= PPG_ST_NUMERIC; n_num++; }
            status[idx] = s2;
            n_lp++; n_piv += res.pivots; n_work += (unsigned long long)res.work;
        }
        if constexpr (NW > 1) __syncthreads();
        }
        group_sync();   // need_list is rewritten by the next block
    }
    if (tid == 0 && n_lp) {
        atomicAdd(&counters[CNT_K2_LPS], n_lp);
        atomicAdd(&counters[CNT_K2_PIVOTS], n_piv);
        atomicAdd(&counters[CNT_K2_WORK], n_work);
        if (n_num) atomicAdd(&counters[CNT_NUMERIC], n_num);
    }
}

template <int NW, int RPT, int DC>
static cudaError_t launch_k2_t(const DevProgram& P, const uint64_t* masks, long long n, uint8_t* status,
                               unsigned long long* queue, unsigned long long* counters, int sm_count, cudaStream_t st) {
    constexpr int WPC = (NW == 1) ? 4 : 1;
    auto kern = k2_feas_kernel<NW, RPT, DC, WPC>;
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32 * WPC, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    long long groups_needed = (n + WPC - 1) / WPC;
    long long grid = (long long)sm_count * occ;
    if (grid > groups_needed) grid = groups_needed;
    if (grid < 1) grid = 1;
    // status bytes scanned per thread and queue item: 16 when K2a has been over the level (next to nothing is left), fewer
    // when the simplex still owns most candidates and blocks must stay small enough to balance
    int scan_pt = 1;
    if (P.warm_count != nullptr) {
        const long long groups = grid * WPC;
        scan_pt = 16;
        while (scan_pt > 1 && (n + (long long)NW * 32 * scan_pt - 1) / ((long long)NW * 32 * scan_pt) < 4 * groups) scan_pt >>= 1;
    }
    kern<<<(unsigned)grid, NW * 32 * WPC, 0, st>>>(P, masks, n, status, queue, counters, scan_pt);
    return cudaGetLastError();
}

#define K2_DC_SWITCH(NW, RPT)                                                                          \
    switch (dc) {                                                                                      \
        case 8: return launch_k2_t<NW, RPT, 8>(P, masks, n, status, queue, counters, sm_count, st);    \
        case 16: return launch_k2_t<NW, RPT, 16>(P, masks, n, status, queue, counters, sm_count, st);  \
        case 24: return launch_k2_t<NW, RPT, 24>(P, masks, n, status, queue, counters, sm_count, st);  \
        case 32: return launch_k2_t<NW, RPT, 32>(P, masks, n, status, queue, counters, sm_count, st);  \
        case 40: return launch_k2_t<NW, RPT, 40>(P, masks, n, status, queue, counters, sm_count, st);  \
        case 48: return launch_k2_t<NW, RPT, 48>(P, masks, n, status, queue, counters, sm_count, st);  \
        case 64: return launch_k2_t<NW, RPT, 64>(P, masks, n, status, queue, counters, sm_count, st);  \
        default: return cudaErrorInvalidValue;                                                         \
    }

#define K2_DC_SWITCH_SMALL(NW, RPT)                                                                    \
    switch (dc) {                                                                                      \
        case 8: return launch_k2_t<NW, RPT, 8>(P, masks, n, status, queue, counters, sm_count, st);    \
        case 16: return launch_k2_t<NW, RPT, 16>(P, masks, n, status, queue, counters, sm_count, st);  \
        case 24: return launch_k2_t<NW, RPT, 24>(P, masks, n, status, queue, counters, sm_count, st);  \
        default: return cudaErrorInvalidValue;                                                         \
    }
#define K2_DC_SWITCH_LARGE(NW, RPT)                                                                    \
    switch (dc) {                                                                                      \
        case 32: return launch_k2_t<NW, RPT, 32>(P, masks, n, status, queue, counters, sm_count, st);  \
        case 40: return launch_k2_t<NW, RPT, 40>(P, masks, n, status, queue, counters, sm_count, st);  \
        case 48: return launch_k2_t<NW, RPT, 48>(P, masks, n, status, queue, counters, sm_count, st);  \
        case 64: return launch_k2_t<NW, RPT, 64>(P, masks, n, status, queue, counters, sm_count, st);  \
        default: return cudaErrorInvalidValue;                                                         \
    }

#define K2_DC_SWITCH_2x2()                                                                           \
    switch (dc) {                                                                                      \
        case 32: return launch_k2_t<2, 2, 32>(P, masks, n, status, queue, counters, sm_count, st);     \
        case 40: return launch_k2_t<2, 2, 40>(P, masks, n, status, queue, counters, sm_count, st);     \
        default: return cudaErrorInvalidValue;                                                         \
    }

__global__ void clear_bits_kernel(uint8_t* __restrict__ status, long long n, uint8_t bits) {
    // 16 status bytes per thread
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i + 16 <= n && (reinterpret_cast<uintptr_t>(status + i) & 15) == 0) {
        uint4 v = *reinterpret_cast<uint4*>(status + i);
        const unsigned m = ~(0x01010101u * bits);
        v.x &= m; v.y &= m; v.z &= m; v.w &= m;
        *reinterpret_cast<uint4*>(status + i) = v;
    } else {
        for (long long j = i; j < n && j < i + 16; ++j) status[j] &= (uint8_t)~bits;
    }
}

cudaError_t launch_clear_bits(uint8_t* status, long long n, uint8_t bits, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const long long threads = (n + 15) / 16;
    clear_bits_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(status, n, bits);
    return cudaGetLastError();
}

int k2_pad_columns(int ncols_with_rhs) {
    const int opts[] = {8, 16, 24, 32, 40, 48, 64};
    for (int o : opts) if (ncols_with_rhs <= o) return o;
    return -1;
}

// chooses the thread mapping from the tableau shape: rows -> warps, columns -> registers per thread
cudaError_t launch_k2(const DevProgram& P, const uint64_t* masks, long long n, uint8_t* status,
                      unsigned long long* queue, unsigned long long* counters, int sm_count, cudaStream_t st) {
    const int dc = k2_pad_columns(P.dc0);
    if (dc < 0 || P.R0 > 256) return cudaErrorInvalidValue;
    if (P.R0 <= 32) { K2_DC_SWITCH(1, 1) }
    if (P.R0 <= 64) {
        if (dc <= 24) { K2_DC_SWITCH_SMALL(1, 2) }
        K2_DC_SWITCH_LARGE(2, 1)
    }
    if (P.R0 <= 128) {
        // 65-128 rows: 2 warps x 2 rows per thread beats 4 warps x 1 row (measured 226 vs 250 ms on the 112 x 38
        // tableau): the per-pivot bookkeeping is paid per warp, the rank-1 update per row.  PPGPU_K2_CFG=4x1 overrides.
        const char* cfg = getenv("PPGPU_K2_CFG");
        if (!(cfg && cfg[0] == '4') && dc <= 40 && dc >= 32) { K2_DC_SWITCH_2x2() }
        K2_DC_SWITCH(4, 1)
    }
    K2_DC_SWITCH(8, 1)
}

}  // namespace ppgpu
