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
        if constexpr (I < DC) v = r[I < DC ? I : 0]; \
        break;
#define PPG_CASES8(B) PPG_CASE(B + 0) PPG_CASE(B + 1) PPG_CASE(B + 2) PPG_CASE(B + 3) PPG_CASE(B + 4) PPG_CASE(B + 5) PPG_CASE(B + 6) PPG_CASE(B + 7)

// r[j] for a register array and a warp-uniform j (a switch keeps the array in registers)
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
        if constexpr (I < DC) r[I < DC ? I : 0] = 0.0; \
        break;
template <int DC>
__device__ __forceinline__ void reg_zero(double (&r)[DC], int j) {
    switch (j) {
        PPG_CASES8(0) PPG_CASES8(8) PPG_CASES8(16) PPG_CASES8(24) PPG_CASES8(32) PPG_CASES8(40) PPG_CASES8(48) PPG_CASES8(56)
        default: break;
    }
}
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

template <int NW, int DC>
struct LpShared {
    double P[DC];           // published pivot row (entry of the pivot column already replaced by 1)
    double pinfo[2];        // effective pivot element, rhs of the pivot row
    double red_v[3][NW];    // cross-warp reduction scratch, one slot per reduction site
    int red_i[3][NW];
    int nbvar[DC];          // variable id of each nonbasic column (Bland's rule)
    unsigned char kind[DC]; // 0 dead, 1 free, 2 slack
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

    // group-wide argmin / argmax / int-min, result known to every thread. `site` selects the scratch slot.
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
    __device__ __forceinline__ static void group_argmax(Shared& sh, int site, int warp, int lane, double& v, int& i) {
        warp_argmax(v, i);
        if constexpr (NW > 1) {
            if (lane == 0) { sh.red_v[site][warp] = v; sh.red_i[site][warp] = i; }
            __syncthreads();
            v = sh.red_v[site][0]; i = sh.red_i[site][0];
#pragma unroll
            for (int w = 1; w < NW; ++w) {
                const double ov = sh.red_v[site][w]; const int oi = sh.red_i[site][w];
                if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
            }
        }
    }
    __device__ __forceinline__ static int group_min_int(Shared& sh, int site, int warp, int lane, int v) {
        v = warp_min_int(v);
        if constexpr (NW > 1) {
            if (lane == 0) sh.red_i[site][warp] = v;
            __syncthreads();
            v = sh.red_i[site][0];
#pragma unroll
            for (int w = 1; w < NW; ++w) v = min(v, sh.red_i[site][w]);
        }
        return v;
    }

    // owner of row r copies it (raw) to shared memory
    __device__ __forceinline__ static void publish_row(Shared& sh, const double (&T)[RPT][DC], int tid, int r) {
        if (tid == r % GT) {
            const int slot = r / GT;
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                if (rr == slot) {
#pragma unroll
                    for (int c = 0; c < DC; ++c) sh.P[c] = T[rr][c];
                }
            });
        }
    }

    // Gauss-Jordan step on column j with the published row sh.P (sh.P[j] must already be 1, inv = 1/pivot).
    // Rows with rflag != 0 other than r are eliminated; row r is rescaled if it stays in the basis.
    __device__ __forceinline__ static void eliminate(const Shared& sh, double (&T)[RPT][DC], const int (&rflag)[RPT], int tid,
                                                     int r, int j, double inv, double dir, bool row_stays) {
        static_for<RPT>([&](auto RR) {
            constexpr int rr = decltype(RR)::value;
            const int row = rr * GT + tid;
            if (row == r) {
                if (row_stays) {
#pragma unroll
                    for (int c = 0; c < DC; ++c) T[rr][c] = sh.P[c] * inv;
                }
            } else if (rflag[rr] != 0) {
                const double f = dir * reg_pick<DC>(T[rr], j) * inv;
                reg_zero<DC>(T[rr], j);
#pragma unroll
                for (int c = 0; c < DC; ++c) T[rr][c] = fma(-f, sh.P[c], T[rr][c]);
            }
        });
    }

    // Solves the LP held in registers. On entry: T rows, rflag (0 dead / 1 inequality / 2 equality), ncol = number
    // of nonbasic columns (columns 1..ncol, the last one, js = ncol, is s), sh.kind/sh.nbvar are initialised here.
    // Returns (uniformly on all threads) once beta >= thr (strict: beta > thr), at optimality, or on failure.
    __device__ static LpOut solve(Shared& sh, double (&T)[RPT][DC], int (&rflag)[RPT], int nrows, int ncol, double thr,
                                  bool strict, int tid) {
        const int lane = tid & 31, warp = tid >> 5;
        const int js = ncol;
        int bvar[RPT];
        static_for<RPT>([&](auto RR) {
            constexpr int rr = decltype(RR)::value;
            bvar[rr] = DC + rr * GT + tid;
        });
        for (int c = tid; c < DC; c += GT) {
            sh.kind[c] = (c >= 1 && c <= ncol) ? 1 : 0;
            sh.nbvar[c] = c;
        }
        LpOut out;
        out.code = PPG_LP_OPTIMAL; out.beta = -CUDART_INF; out.pivots = 0; out.work = 0;
        int live = 0;
        {
            int cnt = 0;
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                cnt += rflag[rr] != 0;
            });
            // live-row count only feeds the work counter; a warp-level sum is enough precision for NW == 1 and
            // for NW > 1 we simply use nrows (upper bound on useful work is not claimed anywhere)
            live = nrows;
            (void)cnt;
        }
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
            publish_row(sh, T, tid, e);
            gsync();
            // largest free coefficient of the row (every warp finds it redundantly)
            double best = 0.0; int j = 0x7fffffff;
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                const int c = cc * 32 + lane;
                if (c < DC && c != js && sh.kind[c] == 1) {
                    const double a = fabs(sh.P[c]);
                    if (a > best) { best = a; j = c; }
                }
            }
            warp_argmax(best, j);
            const double rhs_e = sh.P[0];
            if (j == 0x7fffffff || best < PPG_PIV_TOL) {
                if (fabs(rhs_e) > PPG_FEAS_TOL) { out.code = PPG_LP_INFEAS_EQ; return out; }
                static_for<RPT>([&](auto RR) {
                    constexpr int rr = decltype(RR)::value;
                    if (rr * GT + tid == e) rflag[rr] = 0;
                });
                gsync();
                continue;
            }
            const double piv = sh.P[j];
            gsync();  // everyone has read P[j] before it is overwritten with 1
            if (tid == 0) { sh.P[j] = 1.0; sh.kind[j] = 0; }
            gsync();
            eliminate(sh, T, rflag, tid, e, j, 1.0 / piv, 1.0, false);
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                if (rr * GT + tid == e) rflag[rr] = 0;
            });
            out.pivots++; out.work += (long long)live * (ncol + 1); live--;
            gsync();
        }
        // ---------------- Phase B: s enters the basis on the row of smallest rhs ----------------
        double alpha[CPL];
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
            if (tid == r0 % GT) {
                static_for<RPT>([&](auto RR) {
                    constexpr int rr = decltype(RR)::value;
                    if (rr == r0 / GT) {
#pragma unroll
                        for (int c = 0; c < DC; ++c) sh.P[c] = T[rr][c];
                        sh.pinfo[0] = reg_pick<DC>(T[rr], js);
                    }
                });
                sh.P[js] = 1.0; sh.kind[js] = 2; sh.nbvar[js] = DC + r0;
            }
            gsync();
            const double inv = 1.0 / sh.pinfo[0];
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                const int c = cc * 32 + lane;
                alpha[cc] = (c < DC) ? sh.P[c] * inv : 0.0;
            }
            eliminate(sh, T, rflag, tid, r0, js, inv, 1.0, false);
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                if (rr * GT + tid == r0) rflag[rr] = 0;
            });
            out.pivots++; out.work += (long long)live * (ncol + 1); live--;
            gsync();
        }
        // ---------------- Phase C: primal simplex on the objective "s" ----------------
        int degen = 0; bool bland = false;
        const int cap = 50 * (nrows + ncol) + 200;
        for (int it = 0;; ++it) {
            const double beta = shfl_d(alpha[0], 0);
            out.beta = beta;
            if (strict ? (beta > thr) : (beta >= thr)) { out.code = PPG_LP_EARLY; return out; }
            if (it > cap) { out.code = PPG_LP_ITERLIM; return out; }
            // pricing (lane-parallel over columns, replicated in every warp)
            double best = PPG_OPT_TOL; int j = 0x7fffffff; int bestvar = 0x7fffffff;
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                const int c = cc * 32 + lane;
                if (c >= 1 && c <= ncol) {
                    const int kd = sh.kind[c];
                    double score = -1.0;
                    if (kd == 1) score = fabs(alpha[cc]); else if (kd == 2) score = -alpha[cc];
                    if (score > PPG_OPT_TOL) {
                        if (bland) {
                            const int vid = sh.nbvar[c];
                            if (vid < bestvar) { bestvar = vid; j = c; }
                        } else if (score > best) { best = score; j = c; }
                    }
                }
            }
            if (bland) {
                // smallest variable id wins; carry the column along
                int key = bestvar;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int ok = __shfl_xor_sync(PPG_FULL, key, o);
                    const int oj = __shfl_xor_sync(PPG_FULL, j, o);
                    if (ok < key) { key = ok; j = oj; }
                }
            } else {
                warp_argmax(best, j);
            }
            if (j == 0x7fffffff) { out.code = PPG_LP_OPTIMAL; return out; }
            double aj = 0.0;
#pragma unroll
            for (int cc = 0; cc < CPL; ++cc) {
                const double v = shfl_d(alpha[cc], j & 31);
                if (cc == (j >> 5)) aj = v;
            }
            const bool entering_free = sh.kind[j] == 1;
            const int enter_var = sh.nbvar[j];
            const double dir = (entering_free && aj > 0.0) ? -1.0 : 1.0;
            // ratio test, Harris pass 1: largest admissible step with the rhs relaxed by PPG_HARRIS
            double colv[RPT];
            double tmax = CUDART_INF; int dummy = 0;
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                colv[rr] = dir * reg_pick<DC>(T[rr], j);
                if (rflag[rr] == 1 && colv[rr] > PPG_PIV_TOL) {
                    const double rhs = fmax(T[rr][0], 0.0);
                    tmax = fmin(tmax, (rhs + PPG_HARRIS) / colv[rr]);
                }
            });
            group_argmin(sh, 0, warp, lane, tmax, dummy);
            if (tmax == CUDART_INF) { out.code = PPG_LP_UNBOUNDED; out.beta = CUDART_INF; return out; }
            // pass 2: among rows that block within tmax take the largest pivot (Bland: the smallest basic variable)
            double bp = -1.0; int r = 0x7fffffff;
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                if (rflag[rr] == 1 && colv[rr] > PPG_PIV_TOL) {
                    const double rhs = fmax(T[rr][0], 0.0);
                    if (rhs / colv[rr] <= tmax) {
                        const int row = rr * GT + tid;
                        const double key = bland ? -(double)bvar[rr] : colv[rr];
                        if (key > bp || (key == bp && row < r)) { bp = key; r = row; }
                    }
                }
            });
            group_argmax(sh, 1, warp, lane, bp, r);
            // the owner publishes the pivot row (pivot-column entry := 1), the effective pivot and the row's rhs
            if (tid == r % GT) {
                static_for<RPT>([&](auto RR) {
                    constexpr int rr = decltype(RR)::value;
                    if (rr == r / GT) {
#pragma unroll
                        for (int c = 0; c < DC; ++c) sh.P[c] = T[rr][c];
                        sh.pinfo[0] = colv[rr];
                        sh.pinfo[1] = T[rr][0];
                        sh.nbvar[j] = bvar[rr];
                        if (!entering_free) bvar[rr] = enter_var;
                    }
                });
                sh.P[j] = 1.0;
                sh.kind[j] = 2;
            }
            gsync();
            const double pj = sh.pinfo[0];
            const double step = fmax(sh.pinfo[1], 0.0) / pj;
            const double inv = 1.0 / pj;
            if (step <= PPG_DEGEN_STEP) { if (++degen > PPG_BLAND_AFTER) bland = true; } else degen = 0;
            eliminate(sh, T, rflag, tid, r, j, inv, dir, !entering_free);
            // objective row (replicated): alpha -= (dir*alpha_j*inv) * P', with alpha_j := 0 first
            {
                const double f = dir * aj * inv;
#pragma unroll
                for (int cc = 0; cc < CPL; ++cc) {
                    const int c = cc * 32 + lane;
                    if (c < DC) {
                        const double a0 = (c == j) ? 0.0 : alpha[cc];
                        alpha[cc] = fma(-f, sh.P[c], a0);
                    }
                }
            }
            static_for<RPT>([&](auto RR) {
                constexpr int rr = decltype(RR)::value;
                if (entering_free && rr * GT + tid == r) rflag[rr] = 0;
                if (rflag[rr] == 1 && T[rr][0] < 0.0 && T[rr][0] > -1e-9) T[rr][0] = 0.0;
            });
            out.pivots++; out.work += (long long)live * (ncol + 1);
            if (entering_free) live--;
            gsync();
        }
    }
};

}  // namespace ppgpu
