// Register-resident dense dictionary simplex for  max s  s.t.  a_i.z + g_i s <= h_i  (z free), some rows equalities.
//
// One LP is solved by a GROUP of NW warps (NW == 1: a single warp, no block barriers at all).  Every thread
// owns RPT tableau rows of DC doubles in REGISTERS; the pivot row is the only thing that travels through shared
// memory (DC doubles per pivot), reductions are warp shuffles (+ NW words of shared memory when NW > 1), and the
// objective row is replicated lane-wise so pricing needs no barrier.
//
// This one routine answers the three LP families of the reference's combinatorial path:
//   feasibility   mplp_program.py:439-444       (K2: z = (x,theta), active rows are equalities, feasible iff s* >= -tol)
//   Chebyshev     utils/chebyshev_ball.py:10-63 (K4: z = theta, unit rows, full dimensional iff s* > 1e-8)
//   redundancy    utils/mpqp_utils.py:143-178   (K5: z = theta, one row forced to equality)
// all of which the reference sends one at a time to GLPK/Gurobi through Solver.solve_lp (solver.py:211-246).
//
// Tableau convention (dictionary form): row i reads  w_i = T[i][0] - sum_c T[i][c] * nb_c,  column 0 = rhs.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <type_traits>

#include "tolerances.h"

namespace ppgpu {

#define PPG_FULL 0xffffffffu

#define PPG_CASE(I) \
    case I:         \
        if constexpr (I < DC) asm volatile("mov.f64 %0, %1;" : "=d"(v) : "d"(r[I < DC ? I : 0])); \
        break;
#define PPG_CASES8(B) PPG_CASE(B + 0) PPG_CASE(B + 1) PPG_CASE(B + 2) PPG_CASE(B + 3) PPG_CASE(B + 4) PPG_CASE(B + 5) PPG_CASE(B + 6) PPG_CASE(B + 7)

// r[j] for a register array and a warp-uniform j.  The opaque asm in every case keeps the compiler from turning the
// switch into a 2*DC-long select chain (measured: 285 FSEL / 378 ISETP per pivot in v1); it stays a branch table.
template <int DC>
__device__ __forceinline__ double reg_pick(const double (&r)[DC], int j) {
    double v = 0.0;
    switch (j) {
        PPG_CASES8(0) PPG_CASES8(8) PPG_CASES8(16) PPG_CASES8(24) PPG_CASES8(32) PPG_CASES8(40) PPG_CASES8(48) PPG_CASES8(56)
        default: break;
    }
    return v;
}
#undef PPG_CASE
#define PPG_CASE(I) \
    case I:         \
        if constexpr (I < DC) asm volatile("mov.f64 %0, %1;" : "=d"(r[I < DC ? I : 0]) : "d"(val)); \
        break;
template <int DC>
__device__ __forceinline__ void reg_set(double (&r)[DC], int j, double val) {
    switch (j) {
        PPG_CASES8(0) PPG_CASES8(8) PPG_CASES8(16) PPG_CASES8(24) PPG_CASES8(32) PPG_CASES8(40) PPG_CASES8(48) PPG_CASES8(56)
        default: break;
    }
}
template <int DC>
__device__ __forceinline__ void reg_zero(double (&r)[DC], int j) { reg_set<DC>(r, j, 0.0); }
#undef PPG_CASE
#undef PPG_CASES8

// compile-time loop: guarantees that indices into register arrays are constants (a plain `#pragma unroll`
// is not honoured around the register-select switches above, which would demote the tableau to local memory)
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (N > 0) {
        static_for<N - 1>(f);
        f(std::integral_constant<int, N - 1>{});
    }
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(PPG_FULL, v, src); }
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(PPG_FULL, v, m); }

// (value, index) warp reductions with a deterministic tie-break on the smaller index
__device__ __forceinline__ void warp_argmin(double& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = shfl_xor_d(v, o);
        const int oi = __shfl_xor_sync(PPG_FULL, i, o);
        if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}
__device__ __forceinline__ void warp_argmax(double& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = shfl_xor_d(v, o);
        const int oi = __shfl_xor_sync(PPG_FULL, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}
__device__ __forceinline__ int warp_min_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(PPG_FULL, v, o));
    return v;
}

// min / max of NON-NEGATIVE doubles across a warp with two 32-bit redux.sync each (the IEEE order of non-negative
// doubles is the unsigned order of their bit patterns); ~6 instructions instead of a 5-step shuffle ladder
__device__ __forceinline__ double warp_min_nonneg(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_min_sync(PPG_FULL, hi);
    const unsigned ml = __reduce_min_sync(PPG_FULL, hi == mh ? lo : 0xffffffffu);
    return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ double warp_max_nonneg(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(PPG_FULL, hi);
    const unsigned ml = __reduce_max_sync(PPG_FULL, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}

template <int NW, int DC>
struct LpShared {
    double P[2][NW][DC];    // candidate pivot rows, one per warp, ping-pong over iterations (pivot-column entry := 1)
    double c_ratio[2][NW];  // per-warp ratio-test winner: step length, effective pivot, rhs, row, basic-variable id
    double c_piv[2][NW];
    double c_inv[2][NW];
    double c_hb[2][NW];     // per-warp Harris step bound (INF: the warp has no blocking row)
    double c_rhs[2][NW];
    int c_row[2][NW];
    int c_bvar[2][NW];
    double pinfo[2];        // phase A/B: pivot element
    double red_v[2][NW];    // phase A/B cross-warp reduction scratch
    int red_i[2][NW];
};

struct LpOut {
    int code;
    double beta;
    int pivots;
    long long work;  // sum over pivots of live_rows * columns  (useful FMAs)
};

template <int NW, int RPT, int DC>
struct LpCore {
    static constexpr int GT = NW * 32;
    static constexpr int CPL = (DC + 31) / 32;
    typedef LpShared<NW, DC> Shared;

    __device__ __forceinline__ static void gsync() {
        if constexpr (NW == 1) __syncwarp(); else __syncthreads();
    }

    // group-wide argmin / int-min (phase A/B only), result known to every thread. `site` selects the scratch slot.
    __device__ __forceinline__ static void group_argmin(Shared& sh, int site, int warp, int lane, double& v, int& i) {
        warp_argmin(v, i);
        if constexpr (NW > 1) {
            if (lane == 0) { sh.red_v[site][warp] = v; sh.red_i[site][warp] = i; }
            __syncthreads();
            v = sh.red_v[site][0]; i = sh.red_i[site][0];
#pragma unroll
            for (int w = 1; w < NW; ++w) {
                const double ov = sh.red_v[site][w]; const int oi = sh.red_i[site][w];
                if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
            }
        }
    }
    __device__ __forceinline__ static int group_min_int(Shared& sh, int site, int warp, int lane, int v) {
        v = __reduce_min_sync(PPG_FULL, v);
        if constexpr (NW > 1) {
            if (lane == 0) sh.red_i[site][warp] = v;
            __syncthreads();
            v = sh.red_i[site][0];
#pragma unroll
            for (int w = 1; w < NW; ++w) v = min(v, sh.red_i[site][w]);
        }
        return v;
    }

    // Gauss-Jordan step on column j with the published row P, inv = 1/(dir * pivot), colv[rr] = dir * T[rr][j].
    // The owner publishes P[j] = pivot + 1, so that the plain update  T[j] - f * P[j] = (a_ij - f * a_rj) - f = -f
    // lands the new column (coefficients of the leaving slack) without any per-thread dynamic register write.
    // Rows with rflag != 0 other than r are eliminated; row r is rescaled if it stays in the basis.
    __device__ __forceinline__ static void eliminate(const double* __restrict__ P, double (&T)[RPT][DC], const int (&rflag)[RPT],
                                                     const double (&colv)[RPT], int tid, int r, int j, double inv, bool row_stays) {
        static_for<RPT>([&](auto RR) {
            constexpr int rr = decltype(RR)::value;
            const int row = rr * GT + tid;
            if (row == r) {
                if (row_stays) {
#pragma unroll
                    for (int c = 0; c < DC; ++c) T[rr][c] = P[c] * inv;
                    reg_set<DC>(T[rr], j, inv);
                }
            } else if (rflag[rr] != 0) {
                const double f = colv[rr] * inv;
#pragma unroll
                for (int c = 0; c < DC; ++c) T[rr][c] = fma(-f, P[c], T[rr][c]);
            }
        });
    }

    // Solves the LP held in registers. On entry: T rows, rflag (0 dead / 1 inequality / 2 equality), ncol = number
    // of nonbasic columns (columns 1..ncol, the last one, js = ncol, is s).
    // Returns (uniformly on all threads) once beta >= thr (strict: beta > thr), at optimality, or on failure.
    // keep_free: the row in which a FREE variable becomes basic is kept up to date (rflag 3, never in a ratio test) instead
    // of being dropped, so that the caller can read the solution point afterwards with solution_point() - used where a
    // decision hangs on beta itself: the Harris ratio test and the snapping of tiny negative right-hand sides relax rows
    // by up to 1e-9 per pivot, and on flat polytopes that adds up (2.6e-7 seen on ctrl_alloc_n5 level 5).
    __device__ static LpOut solve(Shared& sh, double (&T)[RPT][DC], int (&rflag)[RPT], int nrows, int ncol, double thr,
                                  bool strict, int tid, bool keep_free = false, int (*bvar_out)[RPT] = nullptr) {
        const int lane = tid & 31, warp = tid >> 5;
        const int js = ncol;
        int bvar[RPT];
        static_for<RPT>([&](auto RR) {
            constexpr int rr = decltype(RR)::value;
            bvar[rr] = DC + rr * GT + tid;
        });
        // column kinds (0 dead, 1 free, 2 slack) and variable ids, replicated lane-wise in every warp
        int kind[CPL], nbv[CPL];
        double alpha[CPL];
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
            const int c = cc * 32 + lane;
            kind[cc] = (c >= 1 && c <= ncol) ? 1 : 0;
            nbv[cc] = c;
            alpha[cc] = 0.0;
        }
        LpOut out;
        out.code = PPG_LP_OPTIMAL; out.beta = -CUDART_INF; out.pivots = 0; out.work = 0;
        int live = nrows;
        double* P0 = &sh.P[0][0][0];
        gsync();
        // ---------------- Phase A: pivot the equality rows out (their slack is fixed at zero) ----------------
        for (;;) {
            int e = 0x7fffffff;
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                if (rflag[rr] == 2) e = min(e, rr * GT + tid);
            });
            e = group_min_int(sh, 0, warp, lane, e);
            if (e == 0x7fffffff) break;
            if (tid == e % GT) {
                static_for<RPT>([&](auto RR) {
                    constexpr int rr = decltype(RR)::value;
                    if (rr == e / GT) {
#pragma unroll
                        for (int c = 0; c < DC; ++c) P0[c] = T[rr][c];
                    }
                });
            }
            gsync();
            // largest free coefficient of the row (every warp finds it redundantly)
            double best = 0.0;
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                const int c = cc * 32 + lane;
                if (c < DC && c != js && kind[cc] == 1) best = fmax(best, fabs(P0[c]));
            }
            const double wbest = warp_max_nonneg(best);
            int j = 0x7fffffff;
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                const int c = cc * 32 + lane;
                if (c < DC && c != js && kind[cc] == 1 && fabs(P0[c]) == wbest) j = min(j, c);
            }
            j = __reduce_min_sync(PPG_FULL, j);
            const double rhs_e = P0[0];
            if (j == 0x7fffffff || wbest < PPG_PIV_TOL) {
                if (fabs(rhs_e) > PPG_FEAS_TOL) { out.code = PPG_LP_INFEAS_EQ; return out; }
                static_for<RPT>([&](auto RR) {
                    constexpr int rr = decltype(RR)::value;
                    if (rr * GT + tid == e) rflag[rr] = 0;
                });
                gsync();
                continue;
            }
            const double piv = P0[j];
            double colv[RPT];
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                colv[rr] = reg_pick<DC>(T[rr], j);
            });
            eliminate(P0, T, rflag, colv, tid, e, j, 1.0 / piv, false);  // column j dies: its entries are never read again
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                if (rr * GT + tid == e) rflag[rr] = 0;
            });
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc)
                if (cc * 32 + lane == j) kind[cc] = 0;
            out.pivots++; out.work += (long long)live * (ncol + 1); live--;
            gsync();
        }
        // ---------------- Phase B: s enters the basis on the row of smallest rhs ----------------
        {
            double mn = CUDART_INF; int r0 = 0x7fffffff;
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                if (rflag[rr] == 1) {
                    const double v = T[rr][0];
                    const int row = rr * GT + tid;
                    if (v < mn || (v == mn && row < r0)) { mn = v; r0 = row; }
                }
            });
            group_argmin(sh, 1, warp, lane, mn, r0);
            if (r0 == 0x7fffffff) { out.code = PPG_LP_UNBOUNDED; out.beta = CUDART_INF; return out; }
            double colv[RPT];
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                colv[rr] = reg_pick<DC>(T[rr], js);
            });
            if (tid == r0 % GT) {
                static_for<RPT>([&](auto RR) {
                    constexpr int rr = decltype(RR)::value;
                    if (rr == r0 / GT) {
#pragma unroll
                        for (int c = 0; c < DC; ++c) P0[c] = T[rr][c];
                        sh.pinfo[0] = colv[rr];
                    }
                });
                P0[js] = sh.pinfo[0] + 1.0;
            }
            gsync();
            const double inv = 1.0 / sh.pinfo[0];
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                const int c = cc * 32 + lane;
                alpha[cc] = (c < DC) ? (c == js ? inv : P0[c] * inv) : 0.0;
                if (c == js) { kind[cc] = 2; nbv[cc] = DC + r0; }
            }
            eliminate(P0, T, rflag, colv, tid, r0, js, inv, false);
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                if (rr * GT + tid == r0) rflag[rr] = 0;
            });
            out.pivots++; out.work += (long long)live * (ncol + 1); live--;
            gsync();
        }
        // ---------------- Phase C: primal simplex on the objective "s", ONE barrier per pivot ----------------
        // Latency-trimmed (profiles/r01: the pivot was a ~800-cycle dependent chain with 2 warps per scheduler):
        //  * Dantzig pricing only needs an approximately largest score -> one 32-bit redux on float keys + ballots;
        //    the optimality test itself stays exact (fp64 compare per lane);
        //  * the ratio test first filters rows with float keys (one redux); the true minimiser is always inside the
        //    filter band, and when the band holds a single row (the common case) no fp64 reduction runs at all;
        //  * the winner publishes 1/pivot, nobody divides after the barrier; beta is kept by every lane.
        int degen = 0; bool bland = false;
        const int cap = 50 * (nrows + ncol) + 200;
        int buf = 0;
        double beta = shfl_d(alpha[0], 0);
        for (int it = 0;; ++it) {
            out.beta = beta;
            if (strict ? (beta > thr) : (beta >= thr)) { out.code = PPG_LP_EARLY; return out; }
            if (it > cap) { out.code = PPG_LP_ITERLIM; return out; }
            // ---- pricing: lanes over columns, identical in every warp (no communication between warps)
            int j;
            {
                double sc[CPL];
                bool el = false;
#pragma unroll
                for (int cc = 0; cc < CPL; ++cc) {
                    sc[cc] = kind[cc] == 1 ? fabs(alpha[cc]) : (kind[cc] == 2 ? -alpha[cc] : 0.0);
                    el = el || sc[cc] > PPG_OPT_TOL;
                }
                if (!__any_sync(PPG_FULL, el)) { out.code = PPG_LP_OPTIMAL; return out; }
                if (!bland) {
                    unsigned key[CPL], lk = 0u;
#pragma unroll
                    for (int cc = 0; cc < CPL; ++cc) {
                        key[cc] = sc[cc] > PPG_OPT_TOL ? __float_as_uint(fmaxf((float)sc[cc], 1e-37f)) : 0u;
                        lk = max(lk, key[cc]);
                    }
                    const unsigned wk = __reduce_max_sync(PPG_FULL, lk);
                    j = 0x7fffffff;
#pragma unroll
                    for (int cc = CPL - 1; cc >= 0; --cc) {
                        const unsigned m = __ballot_sync(PPG_FULL, key[cc] == wk);
                        if (m) j = cc * 32 + __ffs((int)m) - 1;
                    }
                } else {
                    int key = 0x7fffffff;
#pragma unroll
                    for (int cc = 0; cc < CPL; ++cc)
                        if (sc[cc] > PPG_OPT_TOL) key = min(key, nbv[cc]);
                    const int wkey = __reduce_min_sync(PPG_FULL, key);
                    j = 0x7fffffff;
#pragma unroll
                    for (int cc = CPL - 1; cc >= 0; --cc) {
                        const unsigned m = __ballot_sync(PPG_FULL, nbv[cc] == wkey && kind[cc] != 0);
                        if (m) j = cc * 32 + __ffs((int)m) - 1;
                    }
                }
            }
            double aj = 0.0; int kj = 0, enter_var = 0;
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                const double v = shfl_d(alpha[cc], j & 31);
                const int k2 = __shfl_sync(PPG_FULL, kind[cc], j & 31);
                const int n2 = __shfl_sync(PPG_FULL, nbv[cc], j & 31);
                if (cc == (j >> 5)) { aj = v; kj = k2; enter_var = n2; }
            }
            const bool entering_free = kj == 1;
            const double dir = (entering_free && aj > 0.0) ? -1.0 : 1.0;
            // ---- Harris ratio test, one barrier: every warp bounds the step with ITS rows ((rhs + delta) / a, entries down
            // to PPG_TINY take part so that no row can drift by more than delta), pre-selects its largest pivot among
            // the rows blocking within that bound and publishes (ratio, bound, pivot, row); after the barrier the global
            // bound is the smallest published one and the winner is the largest pivot among the candidates that respect
            // it (the warp that owns the global bound always has one).
            double colv[RPT], rat[RPT];
            double lh = CUDART_INF;
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                colv[rr] = dir * reg_pick<DC>(T[rr], j);
                rat[rr] = CUDART_INF;
                if (rflag[rr] == 1 && colv[rr] > PPG_TINY) {
                    const double ivc = 1.0 / colv[rr];
                    const double rhs = T[rr][0] > 0.0 ? T[rr][0] : 0.0;
                    rat[rr] = rhs * ivc;
                    const double hb = (rhs + PPG_HARRIS) * ivc;
                    lh = hb < lh ? hb : lh;
                }
            });
            const double wh = warp_min_nonneg(lh);
            int wrow = 0x7fffffff;
            if (wh != CUDART_INF) {
                double lp = 0.0; int lrow = 0x7fffffff, lb = 0x7fffffff;
                static_for<RPT>([&](auto RR) {
                    constexpr int rr = decltype(RR)::value;
                    if (rat[rr] <= wh) {
                        const int row = rr * GT + tid;
                        const bool better = bland ? (bvar[rr] < lb) : (colv[rr] > lp);
                        if (better) { lp = colv[rr]; lrow = row; lb = bvar[rr]; }
                    }
                });
                if (bland) {
                    const int wb = __reduce_min_sync(PPG_FULL, lb);
                    wrow = __reduce_min_sync(PPG_FULL, (lb == wb) ? lrow : 0x7fffffff);
                } else {
                    const double wp = warp_max_nonneg(lp);
                    wrow = __reduce_min_sync(PPG_FULL, (lrow != 0x7fffffff && lp == wp) ? lrow : 0x7fffffff);
                }
            }
            double* Pw = &sh.P[buf][warp][0];
            if (wrow == 0x7fffffff) {
                if (lane == 0) sh.c_hb[buf][warp] = CUDART_INF;
            } else if (tid == wrow % GT) {
                static_for<RPT>([&](auto RR) {
                    constexpr int rr = decltype(RR)::value;
                    if (rr == wrow / GT) {
#pragma unroll
                        for (int c = 0; c < DC; ++c) Pw[c] = T[rr][c];
                        sh.c_piv[buf][warp] = colv[rr];
                        sh.c_inv[buf][warp] = 1.0 / colv[rr];
                        sh.c_ratio[buf][warp] = rat[rr];
                        sh.c_bvar[buf][warp] = bvar[rr];
                        Pw[j] = dir * colv[rr] + 1.0;
                    }
                });
                sh.c_hb[buf][warp] = wh;
                sh.c_row[buf][warp] = wrow;
            }
            gsync();
            double gt = sh.c_hb[buf][0];
            if constexpr (NW > 1) {
#pragma unroll
                for (int w = 1; w < NW; ++w) { const double h2 = sh.c_hb[buf][w]; gt = h2 < gt ? h2 : gt; }
            }
            if (gt == CUDART_INF) { out.code = PPG_LP_UNBOUNDED; out.beta = CUDART_INF; return out; }
            int gw = 0; double gr = CUDART_INF;
            {
                double gp = -1.0; int gb = 0x7fffffff;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    if (sh.c_hb[buf][w] == CUDART_INF) continue;
                    const double r2 = sh.c_ratio[buf][w];
                    if (!(r2 <= gt)) continue;
                    const double p2 = sh.c_piv[buf][w];
                    const int b2 = sh.c_bvar[buf][w];
                    const bool better = bland ? (b2 < gb) : (p2 > gp);   // ties keep the lower warp = lower row
                    if (better) { gr = r2; gp = p2; gb = b2; gw = w; }
                }
            }
            if (gr == CUDART_INF) { out.code = PPG_LP_UNBOUNDED; out.beta = CUDART_INF; return out; }
            const int r = sh.c_row[buf][gw];
            const int leave_var = sh.c_bvar[buf][gw];
            const double inv = sh.c_inv[buf][gw];
            const double* P = &sh.P[buf][gw][0];
            if (gr <= PPG_DEGEN_STEP) { if (++degen > PPG_BLAND_AFTER) bland = true; } else degen = 0;
            // objective row and column bookkeeping (replicated): alpha -= (alpha_j / pivot) * P'
            {
                const double f = dir * aj * inv;
                beta = fma(-f, P[0], beta);
#pragma unroll
                for (int cc = 0; cc < CPL; ++cc) {
                    const int c = cc * 32 + lane;
                    if (c < DC) {
                        alpha[cc] = fma(-f, P[c], alpha[cc]);
                        if (c == j) { kind[cc] = 2; nbv[cc] = leave_var; }
                    }
                }
            }
            eliminate(P, T, rflag, colv, tid, r, j, inv, !entering_free || keep_free);
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                if (rr * GT + tid == r) {
                    if (entering_free) {
                        rflag[rr] = keep_free ? 3 : 0;
                        // the row now holds dir * x_j (the step variable); remember the sign with the column id
                        if (keep_free && bvar_out) (*bvar_out)[rr] = dir < 0.0 ? -enter_var : enter_var;
                    } else {
                        bvar[rr] = enter_var;
                    }
                }
                if (rflag[rr] == 1 && T[rr][0] < 0.0 && T[rr][0] > -1e-9) T[rr][0] = 0.0;
            });
            out.pivots++; out.work += (long long)live * (ncol + 1);
            if (entering_free) live--;
            buf ^= 1;
            if constexpr (NW == 1) __syncwarp();
        }
    }
};

}  // namespace ppgpu
