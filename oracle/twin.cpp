// CPU checker ("twin") of the batched GPU algorithms.  TEST INFRASTRUCTURE - never linked into the product.
//
// This is NOT the reference's algorithm (that is oracle/ppopt_oracle.py, which follows PPOPT call by call
// with an external LP solver).  It is a sequential statement of what the CUDA kernels compute per candidate
// (same program reduction, same pivoting rules, same thresholds) so that
//   (1) the batched algorithm's DECISIONS can be validated against the reference's golden status bytes on
//       a machine without a GPU, and
//   (2) GPU parity tests can compare the kernels with a second, independent implementation.
// Reference semantics being reproduced per candidate:
//   rank       is_full_rank                   /root/reference/src/ppopt/utils/constraint_utilities.py:222-236
//   feasible   MPLP_Program.check_feasibility /root/reference/src/ppopt/mplp_program.py:411-444
//   optimal    check_optimality               /root/reference/src/ppopt/mpqp_program.py:203-322, mplp_program.py:446-569
//   region     gen_cr_from_active_set(_1d)    /root/reference/src/ppopt/utils/mpqp_utils.py:89-320
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ppopt_b200/csrc/host_math.hpp"
#include "../ppopt_b200/csrc/tolerances.h"

using ppgpu::ReducedProgram;
using ppgpu::vec;

namespace {

struct Lp {
    int nr = 0, nc = 0;        // rows, nonbasic columns (1..nc); column 0 = rhs
    std::vector<double> T;     // nr x (nc+1)
    std::vector<int> rowflag;  // 0 dead, 1 live (basic slack >= 0), 2 pending equality
    std::vector<int> colkind;  // index 1..nc: 0 dead, 1 free, 2 slack
    std::vector<int> bvar, nbvar;
    int js = 0;                // column of the margin variable s
    long pivots = 0;
    double& at(int r, int c) { return T[(size_t)r * (nc + 1) + c]; }
};

struct LpResult { int code; double beta; };

void pivot(Lp& lp, int r, int j, double dir, bool row_stays, std::vector<double>& alpha, bool have_obj) {
    const int nc = lp.nc;
    std::vector<double> P(nc + 1);
    for (int c = 0; c <= nc; ++c) P[c] = lp.at(r, c);
    const double piv = dir * P[j];
    const double inv = 1.0 / piv;
    P[j] = 1.0;
    for (int i = 0; i < lp.nr; ++i) {
        if (i == r || lp.rowflag[i] == 0) continue;
        const double f = dir * lp.at(i, j) * inv;
        lp.at(i, j) = 0.0;
        for (int c = 0; c <= nc; ++c) lp.at(i, c) = std::fma(-f, P[c], lp.at(i, c));
    }
    if (have_obj) {
        const double f = dir * alpha[j] * inv;
        alpha[j] = 0.0;
        for (int c = 0; c <= nc; ++c) alpha[c] = std::fma(-f, P[c], alpha[c]);
    }
    if (row_stays) {
        for (int c = 0; c <= nc; ++c) lp.at(r, c) = P[c] * inv;
    }
    lp.pivots++;
}

// maximise s; returns early once beta >= thr (strict: beta > thr)
LpResult lp_maxmin(Lp& lp, double thr, bool strict) {
    const int nc = lp.nc, nr = lp.nr;
    std::vector<double> alpha(nc + 1, 0.0);
    // Phase A: eliminate pending equality rows
    for (int e = 0; e < nr; ++e) {
        if (lp.rowflag[e] != 2) continue;
        int j = -1; double best = 0.0;
        for (int c = 1; c <= nc; ++c)
            if (lp.colkind[c] == 1 && c != lp.js && std::fabs(lp.at(e, c)) > best) { best = std::fabs(lp.at(e, c)); j = c; }
        if (j < 0 || best < PPG_PIV_TOL) {
            if (std::fabs(lp.at(e, 0)) > PPG_FEAS_TOL) return {PPG_LP_INFEAS_EQ, -INFINITY};
            lp.rowflag[e] = 0;
            continue;
        }
        pivot(lp, e, j, 1.0, false, alpha, false);
        lp.rowflag[e] = 0;
        lp.colkind[j] = 0;
    }
    // Phase B: s enters at the row of minimum rhs
    int r0 = -1; double mn = INFINITY;
    for (int i = 0; i < nr; ++i)
        if (lp.rowflag[i] == 1 && lp.at(i, 0) < mn) { mn = lp.at(i, 0); r0 = i; }
    if (r0 < 0) return {PPG_LP_UNBOUNDED, INFINITY};
    {
        // objective row = new row r0 after the pivot: alpha = P'/piv with P'[js] = 1
        const double piv = lp.at(r0, lp.js);
        for (int c = 0; c <= nc; ++c) alpha[c] = lp.at(r0, c) / piv;
        alpha[lp.js] = 1.0 / piv;
        std::vector<double> dummy;
        pivot(lp, r0, lp.js, 1.0, false, dummy, false);
        lp.rowflag[r0] = 0;
        lp.colkind[lp.js] = 2;
        lp.nbvar[lp.js] = nc + 1 + r0;
    }
    int degen = 0; bool bland = false;
    const long cap = 50L * (nr + nc) + 200;
    for (long it = 0;; ++it) {
        const double beta = alpha[0];
        if (strict ? (beta > thr) : (beta >= thr)) return {PPG_LP_EARLY, beta};
        if (it > cap) return {PPG_LP_ITERLIM, beta};
        // pricing
        int j = -1; double best = PPG_OPT_TOL; int bestvar = 1 << 30;
        for (int c = 1; c <= nc; ++c) {
            double score = -1.0;
            if (lp.colkind[c] == 1) score = std::fabs(alpha[c]);
            else if (lp.colkind[c] == 2) score = -alpha[c];
            if (score <= PPG_OPT_TOL) continue;
            if (bland) { if (lp.nbvar[c] < bestvar) { bestvar = lp.nbvar[c]; j = c; } }
            else if (score > best) { best = score; j = c; }
        }
        if (j < 0) return {PPG_LP_OPTIMAL, beta};
        const double dir = (lp.colkind[j] == 1 && alpha[j] > 0.0) ? -1.0 : 1.0;
        // Harris ratio test. Pass 1: largest step that keeps every basic slack >= -PPG_HARRIS (rows with entries down
        // to PPG_TINY take part: ignoring a 5e-10 entry lets its row drift by 5e-10 x step ~ 1e-7). Pass 2: among the
        // rows that block within that step take the LARGEST pivot (Bland mode: the smallest basic variable id).
        double tmax = INFINITY;
        for (int i = 0; i < nr; ++i) {
            if (lp.rowflag[i] != 1) continue;
            const double a = dir * lp.at(i, j);
            if (a > PPG_TINY) tmax = std::fmin(tmax, (std::fmax(lp.at(i, 0), 0.0) + PPG_HARRIS) / a);
        }
        if (tmax == INFINITY) return {PPG_LP_UNBOUNDED, INFINITY};
        int r = -1; double bpiv = 0.0; int bvar = 1 << 30;
        for (int i = 0; i < nr; ++i) {
            if (lp.rowflag[i] != 1) continue;
            const double a = dir * lp.at(i, j);
            if (a > PPG_TINY && std::fmax(lp.at(i, 0), 0.0) / a <= tmax) {
                if (bland) { if (lp.bvar[i] < bvar) { bvar = lp.bvar[i]; r = i; } }
                else if (a > bpiv) { bpiv = a; r = i; }
            }
        }
        const double step = std::fmax(lp.at(r, 0), 0.0) / (dir * lp.at(r, j));
        if (step <= PPG_DEGEN_STEP) { if (++degen > PPG_BLAND_AFTER) bland = true; } else degen = 0;
        const bool entering_free = lp.colkind[j] == 1;
        const int enter_var = lp.nbvar[j];
        pivot(lp, r, j, dir, !entering_free, alpha, true);
        lp.colkind[j] = 2;
        lp.nbvar[j] = lp.bvar[r];
        if (entering_free) lp.rowflag[r] = 0; else lp.bvar[r] = enter_var;
        for (int i = 0; i < nr; ++i)
            if (lp.rowflag[i] == 1 && lp.at(i, 0) < 0.0 && lp.at(i, 0) > -1e-9) lp.at(i, 0) = 0.0;
    }
}

void lp_init(Lp& lp, int nr, int nfreecols) {
    lp.nr = nr; lp.nc = nfreecols + 1; lp.js = nfreecols + 1;
    lp.T.assign((size_t)nr * (lp.nc + 1), 0.0);
    lp.rowflag.assign(nr, 1);
    lp.colkind.assign(lp.nc + 1, 1);
    lp.colkind[0] = 0;
    lp.bvar.resize(nr); lp.nbvar.resize(lp.nc + 1);
    for (int i = 0; i < nr; ++i) lp.bvar[i] = lp.nc + 1 + i;
    for (int c = 0; c <= lp.nc; ++c) lp.nbvar[c] = c;
    lp.pivots = 0;
}

struct Twin {
    ReducedProgram P;
    long lp_pivots = 0;
};

void active_list(const ReducedProgram& P, const uint64_t* mask, std::vector<int>& act) {
    act.clear();
    for (int i = 0; i < P.mi; ++i)
        if ((mask[i >> 6] >> (i & 63)) & 1ull) act.push_back(i);
}

// K1: column-pivoted Householder QR of the n' x k' matrix whose columns are the active reduced rows
int rank_check(const ReducedProgram& P, const std::vector<int>& act, double* ratio_out) {
    const int k = (int)act.size(), np = P.np;
    if (ratio_out) *ratio_out = 1.0;
    if (k == 0) return 1;
    if (k > np) { if (ratio_out) *ratio_out = 0.0; return 0; }
    std::vector<double> M((size_t)np * k);
    for (int j = 0; j < k; ++j) for (int i = 0; i < np; ++i) M[(size_t)i * k + j] = P.At[(size_t)act[j] * np + i];
    std::vector<char> done(k, 0);
    double first = 0.0, minratio = 1.0;
    for (int s = 0; s < k; ++s) {
        int pj = -1; double best = -1.0;
        for (int j = 0; j < k; ++j) {
            if (done[j]) continue;
            double nn = 0.0;
            for (int i = s; i < np; ++i) nn += M[(size_t)i * k + j] * M[(size_t)i * k + j];
            if (nn > best) { best = nn; pj = j; }
        }
        const double rss = std::sqrt(best);
        if (s == 0) first = rss;
        if (first == 0.0) { if (ratio_out) *ratio_out = 0.0; return 0; }
        const double ratio = rss / first;
        minratio = std::fmin(minratio, ratio);
        if (ratio <= PPG_RANK_TOL) { if (ratio_out) *ratio_out = minratio; return 0; }
        done[pj] = 1;
        // Householder on column pj, rows s..np-1
        std::vector<double> v(np, 0.0);
        for (int i = s; i < np; ++i) v[i] = M[(size_t)i * k + pj];
        const double alpha = v[s] >= 0 ? -rss : rss;
        v[s] -= alpha;
        double vtv = 0.0;
        for (int i = s; i < np; ++i) vtv += v[i] * v[i];
        if (vtv > 0.0) {
            for (int j = 0; j < k; ++j) {
                if (done[j]) continue;
                double d = 0.0;
                for (int i = s; i < np; ++i) d += v[i] * M[(size_t)i * k + j];
                d = 2.0 * d / vtv;
                for (int i = s; i < np; ++i) M[(size_t)i * k + j] -= d * v[i];
            }
        }
    }
    if (ratio_out) *ratio_out = minratio;
    return 1;
}

// K2: feasibility LP  max s : rows + s <= rhs, active rows equalities
int feas_check(Twin& tw, const std::vector<int>& act, double* margin_out, int* code_out) {
    const ReducedProgram& P = tw.P;
    Lp lp;
    lp_init(lp, P.R0, P.nfree);
    const int dc = P.nfree + 2;
    for (int i = 0; i < P.R0; ++i)
        for (int c = 0; c < dc; ++c) lp.at(i, c) = P.T0[(size_t)i * dc + c];
    for (int a : act) { lp.rowflag[a] = 2; lp.at(a, lp.js) = 0.0; }
    LpResult res = lp_maxmin(lp, -PPG_FEAS_TOL, false);
    tw.lp_pivots += lp.pivots;
    if (margin_out) *margin_out = res.beta;
    if (code_out) *code_out = res.code;
    if (res.code == PPG_LP_EARLY || res.code == PPG_LP_UNBOUNDED) return 1;
    if (res.code == PPG_LP_OPTIMAL) return res.beta >= -PPG_FEAS_TOL;
    return 0;
}

// rows a.theta <= f, layout [f | a(t)], returns: 0 not optimal, 1 screen passed; radius in *rad
// zero-row rule and normalisation as in gen_cr_from_active_set (mpqp_utils.py:123-126)
int polytope_test(Twin& tw, std::vector<double>& rows, int nrows, int t, double thr, bool strict, double* rad,
                  std::vector<int>* nonzero_flags, double early_extra = 0.0) {
    std::vector<int> keep;
    bool zero_violation = false;
    for (int i = 0; i < nrows; ++i) {
        double* r = &rows[(size_t)i * (t + 1)];
        double mx = 0.0, nn = 0.0;
        for (int c = 1; c <= t; ++c) { mx = std::fmax(mx, std::fabs(r[c])); nn += r[c] * r[c]; }
        const bool nz = !(mx <= PPG_ZERO_ROW);
        if (nonzero_flags) (*nonzero_flags)[i] = nz;
        if (!nz) { if (r[0] < -PPG_FEAS_TOL) zero_violation = true; continue; }
        const double inv = 1.0 / std::sqrt(nn);
        for (int c = 0; c <= t; ++c) r[c] *= inv;
        keep.push_back(i);
    }
    if (rad) *rad = -INFINITY;
    if (zero_violation) return 0;
    if (t == 1) {
        double mn = -INFINITY, mxv = INFINITY;
        for (int i : keep) {
            const double* r = &rows[(size_t)i * 2];
            if (r[1] > 0) mxv = std::fmin(mxv, r[0] / r[1]); else mn = std::fmax(mn, r[0] / r[1]);
        }
        if (rad) *rad = 0.5 * (mxv - mn);
        return (mn + thr <= mxv) ? 1 : 0;  // thr = PPG_WIDTH_1D (final) or half of it (screen)
    }
    Lp lp;
    lp_init(lp, (int)keep.size(), t);
    for (size_t ii = 0; ii < keep.size(); ++ii) {
        const double* r = &rows[(size_t)keep[ii] * (t + 1)];
        lp.at((int)ii, 0) = r[0];
        for (int c = 1; c <= t; ++c) lp.at((int)ii, c) = r[c];
        lp.at((int)ii, lp.js) = 1.0;
    }
    LpResult res = lp_maxmin(lp, thr + early_extra, strict);   // stops early only beyond thr + early_extra
    tw.lp_pivots += lp.pivots;
    if (rad) *rad = res.beta;
    if (res.code == PPG_LP_EARLY) return 1;
    if (res.code == PPG_LP_OPTIMAL) return strict ? (res.beta > thr) : (res.beta >= thr);
    return 0;  // unbounded Chebyshev LP: the reference's solver returns None -> not full dimensional
}

// K3 (Gram/Cholesky) + K4: theta-space rows from the precomputed Gram, then the polytope test
int kkt_cheb_gram(Twin& tw, const std::vector<int>& act, double* rad) {
    const ReducedProgram& P = tw.P;
    const int k = (int)act.size(), t = P.t, mi = P.mi;
    std::vector<double> S((size_t)k * k), L((size_t)k * (t + 1));
    for (int a = 0; a < k; ++a) for (int b2 = 0; b2 < k; ++b2) S[(size_t)a * k + b2] = P.G[(size_t)act[a] * mi + act[b2]];
    if (k > 0 && !ppgpu::cholesky_lower(S, k)) { if (rad) *rad = -INFINITY; return -1; }
    // Lambda = -S^-1 V_act
    for (int col = 0; col <= t; ++col) {
        std::vector<double> y(k);
        for (int i = 0; i < k; ++i) {
            double s = -P.V[(size_t)act[i] * (t + 1) + col];
            for (int j = 0; j < i; ++j) s -= S[(size_t)i * k + j] * y[j];
            y[i] = s / S[(size_t)i * k + i];
        }
        for (int i = k - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < k; ++j) s -= S[(size_t)j * k + i] * L[(size_t)j * (t + 1) + col];
            L[(size_t)i * (t + 1) + col] = s / S[(size_t)i * k + i];
        }
    }
    for (int j = 0; j < k; ++j) {  // multiplier-sign test over the bounding box of Theta (necessary condition)
        double ub = L[(size_t)j * (t + 1)], mx = 0.0, mag = std::fabs(ub);
        for (int c = 0; c < t; ++c) {
            const double a = L[(size_t)j * (t + 1) + 1 + c];
            mx = std::fmax(mx, std::fabs(a));
            if (a != 0.0) { const double term = std::fmax(a * P.th_lo[c], a * P.th_hi[c]); ub += term; mag += std::fabs(term); }
        }
        if (mx > PPG_ZERO_ROW && ub < -1e-9 * std::fmax(1.0, mag)) { if (rad) *rad = -INFINITY; return 0; }
    }
    std::vector<double> rows((size_t)P.R0 * (t + 1));
    std::vector<int> pos(mi, -1);
    for (int a = 0; a < k; ++a) pos[act[a]] = a;
    for (int i = 0; i < mi; ++i) {
        double* r = &rows[(size_t)i * (t + 1)];
        if (pos[i] >= 0) {
            r[0] = L[(size_t)pos[i] * (t + 1)];
            for (int c = 1; c <= t; ++c) r[c] = -L[(size_t)pos[i] * (t + 1) + c];
        } else {
            for (int c = 0; c <= t; ++c) {
                double s = P.V[(size_t)i * (t + 1) + c];
                for (int a = 0; a < k; ++a) s += P.G[(size_t)i * mi + act[a]] * L[(size_t)a * (t + 1) + c];
                r[c] = c == 0 ? s : -s;
            }
        }
    }
    for (int i = 0; i < P.q; ++i) {
        double* r = &rows[(size_t)(mi + i) * (t + 1)];
        r[0] = P.bt[i];
        for (int c = 1; c <= t; ++c) r[c] = P.At_theta[(size_t)i * t + c - 1];
    }
    const double thr = t == 1 ? 0.5 * PPG_WIDTH_1D : PPG_RADIUS_SCREEN;
    return polytope_test(tw, rows, P.R0, t, thr, false, rad, nullptr);
}

// dense LU with partial pivoting on the KKT matrix of mpqp_program.py:182-190; laws (n+k) x (t+1), [const | theta]
bool kkt_lu(const ReducedProgram& P, const std::vector<int>& act_full, std::vector<double>& laws) {
    const int n = P.n, t = P.t, k = (int)act_full.size(), N = n + k, nrhs = t + 1;
    std::vector<double> M((size_t)N * (N + nrhs), 0.0);
    const int ld = N + nrhs;
    for (int a = 0; a < k; ++a) {
        const int row = act_full[a];
        for (int j = 0; j < n; ++j) M[(size_t)a * ld + j] = P.A[(size_t)row * n + j];
        M[(size_t)a * ld + N] = P.b[row];
        for (int c = 0; c < t; ++c) M[(size_t)a * ld + N + 1 + c] = P.F[(size_t)row * t + c];
    }
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) M[(size_t)(k + i) * ld + j] = P.Q[(size_t)i * n + j];
        for (int a = 0; a < k; ++a) M[(size_t)(k + i) * ld + n + a] = P.A[(size_t)act_full[a] * n + i];
        M[(size_t)(k + i) * ld + N] = -P.c[i];
        for (int c = 0; c < t; ++c) M[(size_t)(k + i) * ld + N + 1 + c] = -P.H[(size_t)i * t + c];
    }
    for (int col = 0; col < N; ++col) {
        int p = col; double best = std::fabs(M[(size_t)col * ld + col]);
        for (int i = col + 1; i < N; ++i)
            if (std::fabs(M[(size_t)i * ld + col]) > best) { best = std::fabs(M[(size_t)i * ld + col]); p = i; }
        if (best == 0.0 || !std::isfinite(best)) return false;
        if (p != col) for (int c = 0; c < ld; ++c) std::swap(M[(size_t)p * ld + c], M[(size_t)col * ld + c]);
        const double inv = 1.0 / M[(size_t)col * ld + col];
        for (int i = col + 1; i < N; ++i) {
            const double f = M[(size_t)i * ld + col] * inv;
            if (f == 0.0) continue;
            for (int c = col + 1; c < ld; ++c) M[(size_t)i * ld + c] -= f * M[(size_t)col * ld + c];
        }
    }
    laws.assign((size_t)N * nrhs, 0.0);
    for (int c = 0; c < nrhs; ++c)
        for (int i = N - 1; i >= 0; --i) {
            double s = M[(size_t)i * ld + N + c];
            for (int j = i + 1; j < N; ++j) s -= M[(size_t)i * ld + j] * laws[(size_t)j * nrhs + c];
            laws[(size_t)i * nrhs + c] = s / M[(size_t)i * ld + i];
        }
    return true;
}

// Region emission for one active set; rows in the reference order [lambda rows, inactive rows, Theta rows].
// flags: bit0 nonzero row, bit1 non-redundant, bit2 exact duplicate of an earlier kept row.
// info[0]=region(0/1) info[1]=radius (t>1) or half width (t==1) info[2]=min bound info[3]=max bound
int emit_region(Twin& tw, const std::vector<int>& act, double* laws_out, double* rows_out, int32_t* flags_out,
                double* info, double* margins_out = nullptr) {
    const ReducedProgram& P = tw.P;
    const int n = P.n, t = P.t, ne = P.ne, m = P.m, kk = (int)act.size();
    std::vector<int> act_full;
    for (int i = 0; i < ne; ++i) act_full.push_back(i);
    for (int a : act) act_full.push_back(ne + a);
    const int k = (int)act_full.size();
    std::vector<double> laws;
    info[0] = 0; info[1] = -INFINITY; info[2] = -INFINITY; info[3] = INFINITY;
    if (!kkt_lu(P, act_full, laws)) return -1;
    for (size_t i = 0; i < laws.size(); ++i) laws_out[i] = laws[i];
    const int nrhs = t + 1, R0 = P.R0;
    std::vector<double> rows((size_t)R0 * nrhs);
    std::vector<char> is_act(m, 0);
    for (int a : act_full) is_act[a] = 1;
    int ri = 0;
    for (int a = 0; a < kk; ++a, ++ri) {  // -C theta <= d for the activated inequalities
        const double* l = &laws[(size_t)(n + ne + a) * nrhs];
        rows[(size_t)ri * nrhs] = l[0];
        for (int c = 1; c <= t; ++c) rows[(size_t)ri * nrhs + c] = -l[c];
    }
    for (int i = 0; i < m; ++i) {
        if (is_act[i]) continue;
        for (int c = 0; c <= t; ++c) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s += P.A[(size_t)i * n + j] * laws[(size_t)j * nrhs + c];
            rows[(size_t)ri * nrhs + c] = c == 0 ? P.b[i] - s : s - P.F[(size_t)i * t + c - 1];
        }
        ++ri;
    }
    for (int i = 0; i < P.q; ++i, ++ri) {
        rows[(size_t)ri * nrhs] = P.bt[i];
        for (int c = 1; c <= t; ++c) rows[(size_t)ri * nrhs + c] = P.At_theta[(size_t)i * t + c - 1];
    }
    std::vector<int> nz(R0, 0);
    double rad = 0.0;
    // t > 1: the LP only stops early beyond PPG_RADIUS_BAND, so that `rad` is the exact radius whenever it matters
    const int ok = polytope_test(tw, rows, R0, t, t == 1 ? PPG_WIDTH_1D : PPG_RADIUS, t != 1, &rad, &nz,
                                 t == 1 ? 0.0 : PPG_RADIUS_BAND);
    for (int i = 0; i < R0; ++i) flags_out[i] = nz[i] ? 1 : 0;
    for (size_t i = 0; i < rows.size(); ++i) rows_out[i] = rows[i];
    info[1] = rad;
    if (!ok) return 0;
    std::vector<int> keep;
    for (int i = 0; i < R0; ++i) if (nz[i]) keep.push_back(i);
    if (t == 1) {
        double mn = -INFINITY, mx = INFINITY;
        for (int i : keep) {
            const double* r = &rows[(size_t)i * 2];
            if (r[1] > 0) mx = std::fmin(mx, r[0] / r[1]); else mn = std::fmax(mn, r[0] / r[1]);
        }
        info[2] = mn; info[3] = mx;
        for (int i : keep) {
            const double* r = &rows[(size_t)i * 2];
            const double v = r[0] / r[1];
            if (mn <= v && v <= mx) flags_out[i] |= 2;
        }
    } else {
        for (size_t a = 0; a < keep.size(); ++a) {  // redundancy LP: row `a` forced to equality
            Lp lp;
            lp_init(lp, (int)keep.size(), t);
            for (size_t ii = 0; ii < keep.size(); ++ii) {
                const double* r = &rows[(size_t)keep[ii] * nrhs];
                lp.at((int)ii, 0) = r[0];
                for (int c = 1; c <= t; ++c) lp.at((int)ii, c) = r[c];
                lp.at((int)ii, lp.js) = ii == a ? 0.0 : 1.0;
            }
            lp.rowflag[a] = 2;
            LpResult res = lp_maxmin(lp, margins_out ? 1e30 : -PPG_REDUND_TOL, false);
            tw.lp_pivots += lp.pivots;
            if (margins_out) margins_out[keep[a]] = res.beta;
            if (margins_out && getenv("TWIN_DEBUG")) fprintf(stderr, "row %d code %d beta %.3e pivots %ld\n", keep[a], res.code, res.beta, lp.pivots);
            bool feas = res.code == PPG_LP_EARLY || res.code == PPG_LP_UNBOUNDED ||
                        (res.code == PPG_LP_OPTIMAL && res.beta >= -PPG_REDUND_TOL);
            if (feas) flags_out[keep[a]] |= 2;
        }
        for (int i = 0; i < R0; ++i) {  // exact duplicates among the kept rows (numpy.unique semantics)
            if ((flags_out[i] & 3) != 3) continue;
            for (int j = 0; j < i; ++j) {
                if ((flags_out[j] & 3) != 3 || (flags_out[j] & 4)) continue;
                bool same = true;  // value equality (so -0.0 == 0.0), as numpy.unique compares
                for (int c = 0; c < nrhs && same; ++c) same = rows[(size_t)i * nrhs + c] == rows[(size_t)j * nrhs + c];
                if (same) { flags_out[i] |= 4; break; }
            }
        }
    }
    info[0] = 1;
    return 1;
}


// ---- K2a: sequential statement of k2a_relax_reg_kernel (ppopt_b200/csrc/k2a_relax.cu) -------------------------------
// Same algorithm and constants (LDL' of Gam[A,A], residual recursion through M = Gam[:,A] S^-1, 32-bit arg-max keys, the two
// relaxation phases with their stall detectors, step log, exact verification in Gram space).  The step length uses an
// exact reciprocal where the kernel uses MUFU.RCP64H, so step COUNTS may differ by a few; what the CPU tests pin is the
// property the product relies on: a certified candidate is feasible, and the residuals handed to the simplex are exact.
struct K2aOut { int certified; int steps; std::vector<double> resid; };

static inline int hi_word(double v) { int64_t b; std::memcpy(&b, &v, 8); return (int)(b >> 32); }
static inline double from_hi(int hi) { int64_t b = (int64_t)(uint32_t)hi << 32; double v; std::memcpy(&v, &b, 8); return v; }

K2aOut k2a_certify(const ReducedProgram& P, const std::vector<int>& act, int max_iter, int max_iter2) {
    const int R0 = P.R0, k = (int)act.size(), dc = P.nfree + 2;
    K2aOut out{0, 0, {}};
    auto G = [&](int i, int j) { return P.Gam[(size_t)i * R0 + j]; };
    auto h = [&](int r) { return P.T0[(size_t)r * dc]; };
    // S = L D L'
    std::vector<double> L((size_t)k * k, 0.0), dd(k), dinv(k);
    for (int i = 0; i < k; ++i) for (int j = 0; j <= i; ++j) { if (i == j) dd[i] = G(act[i], act[j]); else L[(size_t)i * k + j] = G(act[i], act[j]); }
    for (int j = 0; j < k; ++j) {
        std::vector<double> wj(k, 0.0);
        for (int c = 0; c < j; ++c) { wj[c] = L[(size_t)j * k + c] * dd[c]; dd[j] = std::fma(-L[(size_t)j * k + c], wj[c], dd[j]); }
        if (!(dd[j] > 1e-300)) return out;
        dinv[j] = 1.0 / dd[j];
        for (int i = j + 1; i < k; ++i) {
            double e = L[(size_t)i * k + j];
            for (int c = 0; c < j; ++c) e = std::fma(-L[(size_t)i * k + c], wj[c], e);
            L[(size_t)i * k + j] = e * dinv[j];
        }
    }
    auto solve = [&](std::vector<double>& x) {
        for (int i = 1; i < k; ++i) for (int c = 0; c < i; ++c) x[i] = std::fma(-L[(size_t)i * k + c], x[c], x[i]);
        for (int i = 0; i < k; ++i) x[i] *= dinv[i];
        for (int i = k - 2; i >= 0; --i) for (int c = i + 1; c < k; ++c) x[i] = std::fma(-L[(size_t)c * k + i], x[c], x[i]);
    };
    std::vector<double> w0(k);
    for (int a = 0; a < k; ++a) w0[a] = h(act[a]);
    solve(w0);
    std::vector<char> isa(R0, 0);
    for (int a : act) isa[a] = 1;
    std::vector<double> M((size_t)R0 * k), v(R0);
    for (int r = 0; r < R0; ++r) {
        std::vector<double> x(k);
        double s2 = -h(r);
        for (int a = 0; a < k; ++a) { x[a] = G(act[a], r); s2 = std::fma(x[a], w0[a], s2); }
        solve(x);
        for (int a = 0; a < k; ++a) M[(size_t)r * k + a] = x[a];
        v[r] = isa[r] ? -1e300 : s2;
    }
    std::vector<double> log_tau;
    std::vector<int> log_row;
    auto exact = [&](std::vector<double>& s) {
        s.assign(R0, 0.0);
        std::vector<double> t(k, 0.0);
        for (int r = 0; r < R0; ++r) s[r] = -h(r);
        for (size_t e = 0; e < log_tau.size(); ++e) {
            for (int r = 0; r < R0; ++r) s[r] = std::fma(-log_tau[e], G(log_row[e], r), s[r]);
            for (int a = 0; a < k; ++a) t[a] = std::fma(-log_tau[e], G(log_row[e], act[a]), t[a]);
        }
        std::vector<double> x(k);
        for (int a = 0; a < k; ++a) x[a] = h(act[a]) - t[a];
        solve(x);
        for (int a = 0; a < k; ++a) for (int r = 0; r < R0; ++r) s[r] = std::fma(x[a], G(act[a], r), s[r]);
    };
    const int ktol = (hi_word(PPG_FEAS_TOL * 0.999) & ~127) | 127;
    if (max_iter > 192) max_iter = 192;
    if (max_iter + max_iter2 > 192) max_iter2 = 192 - max_iter;
    double omega = 1.35;
    int it_end = max_iter, chk = 0, rechecks = 0, wref = 0;
    bool second = max_iter2 <= 0, feasible = false;
    std::vector<double> s;
    for (int it = 0;; ++it) {
        int wkey = 0;
        for (int r = 0; r < R0; ++r) wkey = std::max(wkey, (hi_word(v[r]) & ~127) | r);
        if (wkey <= ktol) {
            exact(s);
            double worst = 0.0;
            for (int r = 0; r < R0; ++r) {
                worst = std::fmax(worst, isa[r] ? std::fabs(s[r]) : s[r]);
                v[r] = isa[r] ? -1e300 : s[r];
            }
            if (worst <= PPG_FEAS_TOL) { feasible = true; break; }
            if (++rechecks > 3) break;
            continue;
        }
        const int irow = wkey & 127;
        const double wmax = from_hi(wkey & ~127);
        bool stalled = it >= it_end;
        if (!stalled && chk == 0) {
            stalled = it != 0 && wkey > wref - 0x100000;
            wref = wkey;
            chk = second ? 48 : 16;
        }
        if (stalled) {
            if (second) break;
            second = true; omega = 1.8; it_end = it + max_iter2; wref = wkey; chk = 48;
        }
        --chk;
        std::vector<double> g2(k), c2(R0);
        for (int a = 0; a < k; ++a) g2[a] = G(act[a], irow);
        for (int r = 0; r < R0; ++r) {
            double x2 = G(irow, r);
            for (int a = 0; a < k; ++a) x2 = std::fma(-M[(size_t)r * k + a], g2[a], x2);
            c2[r] = x2;
        }
        const double nn = c2[irow];
        if (!(nn > 1e-12)) break;
        const double tau = (omega * wmax) * (1.0 / nn);
        log_tau.push_back(tau); log_row.push_back(irow);
        for (int r = 0; r < R0; ++r) v[r] = std::fma(-tau, c2[r], v[r]);
    }
    out.certified = feasible ? 1 : 0;
    out.steps = (int)log_tau.size();
    if (!feasible) { exact(s); out.resid = s; }
    return out;
}


// ---- K2a, prefix form: sequential statement of k2p_relax_kernel (ppopt_b200/csrc/k2p_prefix.cu) -----------------------
// Gram matrix projected once per PREFIX (all active rows but the last two) through the numeric W = S_P^-1 Gam[P,:], the
// candidate's own two rows deflated by a 2x2 solve, point carried as a coefficient vector over the generators, exact
// verification of every row from the coefficients.  Same constants and stall logic as the kernel; the step length uses
// an exact reciprocal (the kernel: MUFU.RCP64H), so step counts may differ by a few.
K2aOut k2p_certify(const ReducedProgram& P, const std::vector<int>& act, int max_iter, int max_iter2) {
    const int R0 = P.R0, k = (int)act.size(), dc = P.nfree + 2;
    K2aOut out{0, 0, {}};
    if (k < 1) return out;
    const int p = k >= 2 ? k - 2 : 0;
    const bool has_b = k >= 2;
    auto G = [&](int i, int j) { return P.Gam[(size_t)i * R0 + j]; };
    auto h = [&](int r) { return P.T0[(size_t)r * dc]; };
    // LDL' of S_P, W, w0
    std::vector<double> Sp((size_t)p * p), Wt((size_t)p * R0), w0(p);
    for (int a = 0; a < p; ++a) for (int b = 0; b < p; ++b) Sp[(size_t)a * p + b] = G(act[a], act[b]);
    for (int j = 0; j < p; ++j) {
        const double d = Sp[(size_t)j * p + j];
        if (!(d > 1e-12)) return out;
        for (int i = j + 1; i < p; ++i) {
            const double lij = Sp[(size_t)i * p + j] / d;
            for (int c = j + 1; c <= i; ++c) Sp[(size_t)i * p + c] = std::fma(-lij, Sp[(size_t)c * p + j], Sp[(size_t)i * p + c]);
            Sp[(size_t)i * p + j] = lij;
        }
    }
    auto solveP = [&](auto get, auto set) {
        for (int i = 1; i < p; ++i) { double x = get(i); for (int c = 0; c < i; ++c) x = std::fma(-Sp[(size_t)i * p + c], get(c), x); set(i, x); }
        for (int i = 0; i < p; ++i) set(i, get(i) / Sp[(size_t)i * p + i]);
        for (int i = p - 2; i >= 0; --i) { double x = get(i); for (int c = i + 1; c < p; ++c) x = std::fma(-Sp[(size_t)c * p + i], get(c), x); set(i, x); }
    };
    for (int j = 0; j < R0; ++j) {
        for (int a = 0; a < p; ++a) Wt[(size_t)a * R0 + j] = G(act[a], j);
        solveP([&](int i) { return Wt[(size_t)i * R0 + j]; }, [&](int i, double x) { Wt[(size_t)i * R0 + j] = x; });
    }
    for (int a = 0; a < p; ++a) w0[a] = h(act[a]);
    solveP([&](int i) { return w0[i]; }, [&](int i, double x) { w0[i] = x; });
    std::vector<double> Gp((size_t)R0 * R0), v0(R0);
    for (int j = 0; j < R0; ++j)
        for (int r = 0; r < R0; ++r) {
            double x = G(j, r);
            for (int a = 0; a < p; ++a) x = std::fma(-G(act[a], r), Wt[(size_t)a * R0 + j], x);
            Gp[(size_t)j * R0 + r] = x;
        }
    for (int r = 0; r < R0; ++r) {
        double x = -h(r);
        for (int a = 0; a < p; ++a) x = std::fma(G(act[a], r), w0[a], x);
        v0[r] = x;
    }
    const int b_row = act[k - 1], a_row = has_b ? act[k - 2] : b_row;
    auto gp = [&](int j, int r) { return Gp[(size_t)j * R0 + r]; };
    const double s11 = gp(a_row, a_row), s12 = has_b ? gp(a_row, b_row) : 0.0, s22 = has_b ? gp(b_row, b_row) : 1.0;
    const double det = std::fma(s11, s22, -s12 * s12);
    if (!(s11 > 1e-12 && s22 > 1e-12 && det > 1e-12 * s11 * s22)) return out;
    const double dinv = 1.0 / det;
    const double i11 = s22 * dinv, i12 = -s12 * dinv, i22 = has_b ? s11 * dinv : 0.0;
    std::vector<char> park(R0, 0);
    for (int a : act) park[a] = 1;
    std::vector<double> v(R0), coef(R0, 0.0), s(R0);
    {
        const double va = v0[a_row], vb = has_b ? v0[b_row] : 0.0;
        const double xa = std::fma(i11, va, i12 * vb), xb = std::fma(i12, va, i22 * vb);
        for (int r = 0; r < R0; ++r) {
            const double cb = has_b ? gp(b_row, r) : 0.0;
            v[r] = park[r] ? -1e300 : std::fma(-xb, cb, std::fma(-xa, gp(a_row, r), v0[r]));
        }
    }
    auto exact = [&]() {
        for (int r = 0; r < R0; ++r) s[r] = v0[r];
        double ua = v0[a_row], ub = has_b ? v0[b_row] : 0.0;
        for (int j = 0; j < R0; ++j) {
            if (coef[j] == 0.0) continue;
            for (int r = 0; r < R0; ++r) s[r] = std::fma(coef[j], gp(j, r), s[r]);
            ua = std::fma(coef[j], gp(j, a_row), ua);
            ub = std::fma(coef[j], gp(j, b_row), ub);
        }
        if (!has_b) ub = 0.0;
        const double xa = -std::fma(i11, ua, i12 * ub), xb = -std::fma(i12, ua, i22 * ub);
        for (int r = 0; r < R0; ++r) s[r] = std::fma(xa, gp(a_row, r), std::fma(xb, has_b ? gp(b_row, r) : 0.0, s[r]));
    };
    const int ktol = (hi_word(PPG_FEAS_TOL * 0.999) & ~127) | 127;
    double omega = 1.35;
    int left = max_iter, chk = 0, wref = 0, rechecks = 0, nst = 0;
    bool second = max_iter2 <= 0, first = true, feasible = false;
    for (;;) {
        int wkey = 0;
        for (int r = 0; r < R0; ++r) wkey = std::max(wkey, (hi_word(v[r]) & ~127) | r);
        int mode = 0;
        if (wkey <= ktol) {
            mode = 1;
        } else if (std::min(chk, left) <= 0) {
            bool stalled = left <= 0;
            if (!stalled) { stalled = !first && wkey > wref - 0x100000; wref = wkey; chk = second ? 48 : 16; }
            if (stalled) {
                if (second) mode = 2;
                else { second = true; omega = 1.8; left = max_iter2; wref = wkey; chk = 48; }
            }
        }
        first = false; --chk; --left;
        if (mode == 0) {
            const int j = wkey & 127;
            const double ga = gp(j, a_row), gb = gp(j, b_row);
            const double m0 = std::fma(ga, i11, gb * i12), m1 = std::fma(ga, i12, gb * i22);
            const double nn = std::fma(-m1, gb, std::fma(-m0, ga, gp(j, j)));
            if (!(nn > 1e-12)) {
                mode = 2;
            } else {
                const double tau = (omega * from_hi(wkey & ~127)) * (1.0 / nn);
                const double t0 = tau * m0, t1 = tau * m1;
                for (int r = 0; r < R0; ++r) {
                    const double cb = has_b ? gp(b_row, r) : 0.0;
                    v[r] = std::fma(t1, cb, std::fma(t0, gp(a_row, r), std::fma(-tau, gp(j, r), v[r])));
                }
                coef[j] -= tau;
                ++nst;
            }
        }
        if (mode != 0) {
            exact();
            double worst = 0.0;
            for (int r = 0; r < R0; ++r) worst = std::fmax(worst, park[r] ? std::fabs(s[r]) : s[r]);
            double eqm = 0.0, sm = 1.0;
            for (int r = 0; r < R0; ++r) { sm = std::fmax(sm, std::fabs(s[r])); if (park[r]) eqm = std::fmax(eqm, std::fabs(s[r])); }
            if (mode == 1 && worst <= PPG_FEAS_TOL && eqm <= 1e-9 * sm) { feasible = true; break; }
            if (mode == 1 && rechecks < 3) {
                ++rechecks;
                for (int r = 0; r < R0; ++r) v[r] = park[r] ? -1e300 : s[r];
                continue;
            }
            break;
        }
    }
    out.certified = feasible ? 1 : 0;
    out.steps = nst;
    if (!feasible) {
        // the simplex may only start from this point if its residual vector is trustworthy: on an ill-conditioned prefix
        // the projector W is large, the rounding of Gp (eps |W|) shows up as non-zero residuals on the active rows, and
        // every other row carries the same noise - such candidates go to the simplex cold (resid[0] = NaN)
        double eqmax = 0.0, smax = 1.0;
        for (int r = 0; r < R0; ++r) { smax = std::fmax(smax, std::fabs(s[r])); if (park[r]) eqmax = std::fmax(eqmax, std::fabs(s[r])); }
        out.resid = s;
        if (!(eqmax <= 1e-9 * smax)) out.resid[0] = NAN;
    }
    return out;
}

// ---- K2w: sequential statement of k2w_walk_kernel (ppopt_b200/csrc/k2w_walk.cu) -------------------------------------------
// One walker processes the candidates of a level in order, prefix by prefix (a prefix = all active rows but the last two),
// exactly like a device walker processes one work item: slack dictionary s_B = beta - D s_N reloaded from the host-built
// vertex, drive(row) by primal simplex pivots (largest positive coefficient of the target row enters, Harris ratio test,
// the target row wins when eligible), prefix rows fixed first, then for every second-last row a: drive a, fix it, drive
// every still-open b; after every pivot all open candidates whose last two rows are nonbasic are certified.  Same
// constants as the kernel (pivot floor 1e-7, certification band 1e-8, 400 pivots per drive, reload after 3000 pivots).
struct WalkDict {
    int nb, nf, ld;
    std::vector<double> D;        // nb x ld  [beta | coefficients]
    std::vector<int> bvar, nvar, where;
    long pivots = 0;
    bool ok = true;               // every basic slack >= -1e-8 after the last pivot
    void load(const ReducedProgram& P) {
        nb = P.wk_nb; nf = P.nfree; ld = P.wk_ld;
        D = P.wk_D0; bvar = P.wk_bvar; nvar = P.wk_nvar;
        where.assign(P.R0, 0);
        for (int i = 0; i < nb; ++i) where[bvar[i]] = i;
        for (int j = 0; j < nf; ++j) where[nvar[j]] = ~j;
        ok = true;
    }
    bool nonbasic(int r) const { return where[r] < 0; }
    void pivot(int l, int j) {
        const int cj = 1 + j;
        const double inv = 1.0 / D[(size_t)l * ld + cj];
        std::vector<double> q(ld);
        for (int k = 0; k < ld; ++k) q[k] = k == cj ? 0.0 : D[(size_t)l * ld + k] * inv;
        ok = true;
        for (int r = 0; r < nb; ++r) {
            if (r == l) continue;
            const double col = D[(size_t)r * ld + cj];
            if (col != 0.0) {
                for (int k = 0; k < ld; ++k) D[(size_t)r * ld + k] = std::fma(-col, q[k], D[(size_t)r * ld + k]);
                D[(size_t)r * ld + cj] = -col * inv;
            }
            if (!(D[(size_t)r * ld] >= -1e-8)) ok = false;
        }
        for (int k = 0; k < ld; ++k) D[(size_t)l * ld + k] = k == cj ? inv : q[k];
        const int rl = bvar[l], rn = nvar[j];
        bvar[l] = rn; nvar[j] = rl; where[rn] = l; where[rl] = ~j;
        ++pivots;
    }
    // returns the row that became nonbasic by the pivot (or -1 when no pivot was possible), reached = target is nonbasic
    template <class OnPivot>
    bool drive(int t, const std::vector<char>& fixed_col, OnPivot&& on_pivot) {
        for (int it = 0; it < 400; ++it) {
            const int wi = where[t];
            if (wi < 0) return true;
            const double bi = D[(size_t)wi * ld];
            const bool degen = bi <= 1e-11;
            double best = 0.0; int j = -1;
            for (int jj = 0; jj < nf; ++jj) {
                if (fixed_col[jj]) continue;
                double x = D[(size_t)wi * ld + 1 + jj];
                if (degen) x = std::fabs(x);
                if (x > best) { best = x; j = jj; }
            }
            if (!(best > 1e-7)) return false;
            int l = wi;
            if (!degen) {
                double hb = INFINITY;
                for (int r = 0; r < nb; ++r) {
                    const double a = D[(size_t)r * ld + 1 + j];
                    if (a > PPG_TINY) hb = std::fmin(hb, (std::fmax(D[(size_t)r * ld], 0.0) + PPG_HARRIS) / a);
                }
                const double at = D[(size_t)wi * ld + 1 + j];
                if (!(std::fmax(bi, 0.0) / at <= hb)) {
                    double lp = 0.0; l = -1;
                    for (int r = 0; r < nb; ++r) {
                        const double a = D[(size_t)r * ld + 1 + j];
                        if (a > PPG_TINY && std::fmax(D[(size_t)r * ld], 0.0) / a <= hb && a > lp) { lp = a; l = r; }
                    }
                    if (l < 0) return false;
                }
            }
            const int rl = bvar[l];
            pivot(l, j);
            if (!ok) return false;
            on_pivot(rl);
        }
        return false;
    }
};

// certified[i] = 1 for every candidate the walk certifies; returns the number of pivots.
// witness (n x W4 words, optional): the mask of ALL rows nonbasic at the vertex that certified candidate i - what the kernel
// leaves for the next level (k6_children.cu::inherit_kernel: a child one of whose parents has a witness holding the added
// row too is certified by that same vertex).  W4 = ceil(R0 / 64).
// slots >= 2: slot 1 receives a LATER vertex of the walk that holds an already closed candidate as well (the kernel's
// k2w_revisit_row / k2w_revisit_all: after a pivot that brought row r in, every closed candidate {r, y} with y nonbasic;
// after the prefix has been fixed, every closed candidate with both rows nonbasic).  closed (optional, n flags): candidates
// that arrive certified (by inheritance): part of the segment, never a target, revisited like the others.
long k2w_walk_level(const ReducedProgram& P, const uint64_t* masks, long n, std::vector<char>& certified,
                    uint64_t* witness = nullptr, int slots = 1, const uint8_t* closed = nullptr) {
    certified.assign(n, 0);
    if (!P.wk_ok || n == 0) return 0;
    std::vector<std::vector<int>> acts(n);
    for (long i = 0; i < n; ++i) active_list(P, masks + i * P.W, acts[i]);
    const int k = (int)acts[0].size();
    if (k < 1) return 0;
    const int p = k >= 2 ? k - 2 : 0;
    WalkDict wd;
    wd.load(P);
    long total = 0;
    std::vector<int> fixed_rows;
    long i = 0;
    while (i < n) {
        long s1 = i + 1;
        while (s1 < n && std::equal(acts[i].begin(), acts[i].begin() + p, acts[s1].begin())) ++s1;
        // open candidates of the segment: (a, b) -> index
        std::vector<std::vector<long>> open(P.R0, std::vector<long>(P.R0, -1)), orig;
        for (long q = i; q < s1; ++q) {
            const int b = acts[q][k - 1], a = k >= 2 ? acts[q][k - 2] : b;
            open[a][b] = q;
        }
        orig = open;
        if (closed)
            for (long q = i; q < s1; ++q)
                if (closed[q]) open[k >= 2 ? acts[q][k - 2] : acts[q][k - 1]][acts[q][k - 1]] = -1;
        const int W4 = (P.R0 + 63) / 64;
        bool second_on = true;   // off while k2w_mark_all is restated: it writes first witnesses only
        auto write_witness = [&](long q, int slot) {
            uint64_t* wq = witness + ((size_t)q * slots + slot) * W4;
            for (int w = 0; w < W4; ++w) wq[w] = 0;
            for (int j = 0; j < wd.nf; ++j) wq[wd.nvar[j] >> 6] |= 1ull << (wd.nvar[j] & 63);
        };
        auto mark_row = [&](int r) {
            for (int y = 0; y < P.R0; ++y) {
                if (!wd.nonbasic(y)) continue;
                const int a = std::min(r, y), b = std::max(r, y);
                if (open[a][b] >= 0) {
                    certified[open[a][b]] = 1;
                    if (witness) write_witness(open[a][b], 0);
                    open[a][b] = -1;
                }
                if (witness && second_on && slots >= 2 && k >= 2 && y != r && orig[a][b] >= 0 && open[a][b] < 0) write_witness(orig[a][b], 1);
            }
        };
        auto revisit_all = [&]() {
            if (!witness || slots < 2 || k < 2) return;
            for (int a = 0; a < P.R0; ++a)
                for (int b = a + 1; b < P.R0; ++b)
                    if (wd.nonbasic(a) && wd.nonbasic(b) && orig[a][b] >= 0 && open[a][b] < 0) write_witness(orig[a][b], 1);
        };
        bool restart = wd.pivots > 3000 || !wd.ok;
        auto fix_prefix = [&]() -> bool {
            if (restart) { total += wd.pivots; wd.pivots = 0; wd.load(P); fixed_rows.clear(); restart = false; }
            size_t keep = 0;
            while (keep < fixed_rows.size() && keep < (size_t)p && fixed_rows[keep] == acts[i][keep]) ++keep;
            fixed_rows.resize(keep);
            for (int f = (int)keep; f < p; ++f) {
                std::vector<char> fc(wd.nf, 0);
                for (int r : fixed_rows) fc[~wd.where[r]] = 1;
                if (!wd.drive(acts[i][f], fc, [](int) {})) return false;
                fixed_rows.push_back(acts[i][f]);
            }
            return true;
        };
        if (fix_prefix()) {
            // (the kernel's k2w_mark_all writes first witnesses only; second ones come from k2w_revisit_all right after it)
            { second_on = false; for (int r = 0; r < P.R0; ++r) if (wd.nonbasic(r)) mark_row(r); second_on = true; }
            revisit_all();
            for (int a = 0; a < P.R0 && wd.ok; ++a) {
                bool any = false;
                for (int b = a; b < P.R0; ++b) any = any || open[a][b] >= 0;
                if (!any) continue;
                if (wd.pivots > 3000) {
                    restart = true;
                    if (!fix_prefix()) break;
                    { second_on = false; for (int r = 0; r < P.R0; ++r) if (wd.nonbasic(r)) mark_row(r); second_on = true; }
                    revisit_all();
                }
                std::vector<char> fc(wd.nf, 0);
                for (int r : fixed_rows) fc[~wd.where[r]] = 1;
                if (!wd.drive(a, fc, mark_row)) { for (int b = a; b < P.R0; ++b) open[a][b] = -1; continue; }
                if (k < 2) continue;
                fc[~wd.where[a]] = 1;
                for (int b = a + 1; b < P.R0 && wd.ok; ++b) {
                    if (open[a][b] < 0) continue;
                    wd.drive(b, fc, mark_row);
                    open[a][b] = -1;   // certified by the marking, or left to the relaxation
                }
            }
        }
        i = s1;
    }
    return total + wd.pivots;
}

}  // namespace

extern "C" {

void* twin_create(int n, int t, int m, int q, int ne, int is_qp, const double* A, const double* b, const double* F,
                  const double* A_t, const double* b_t, const double* Q, const double* c, const double* H) {
    Twin* tw = new Twin();
    if (!ppgpu::reduce_program(n, t, m, q, ne, is_qp, A, b, F, A_t, b_t, Q, c, H, tw->P)) { delete tw; return nullptr; }
    return tw;
}
void twin_destroy(void* h) { delete (Twin*)h; }
int twin_words(void* h) { return ((Twin*)h)->P.W; }
int twin_use_gram(void* h) { return ((Twin*)h)->P.use_gram; }
long twin_pivots(void* h) { return ((Twin*)h)->lp_pivots; }

// status byte per candidate; aux (n x 3): [rank ratio, feasibility margin, radius]
// screen_only=1: stop at the Gram screen (bit PPG_ST_OPT); 0: also run the LU emission test (bit PPG_ST_REGION)
void twin_eval(void* h, const uint64_t* masks, long ncand, int final_level_lp, uint8_t* status, double* aux) {
    Twin& tw = *(Twin*)h;
    const ReducedProgram& P = tw.P;
    std::vector<int> act;
    std::vector<double> laws((size_t)(P.n + P.m) * (P.t + 1)), rows((size_t)P.R0 * (P.t + 1));
    std::vector<int32_t> flags(P.R0);
    for (long ci = 0; ci < ncand; ++ci) {
        active_list(P, masks + ci * P.W, act);
        uint8_t st = 0;
        double ratio = 1.0, margin = 0.0, rad = -INFINITY;
        if (rank_check(P, act, &ratio)) st |= PPG_ST_RANK;
        if (ratio > PPG_RANK_BORDER_LO && ratio < PPG_RANK_BORDER_HI) st |= PPG_ST_BORDER;
        if (st & PPG_ST_RANK) {
            int code = 0;
            if (feas_check(tw, act, &margin, &code)) st |= PPG_ST_FEAS;
            if (code == PPG_LP_ITERLIM) st |= PPG_ST_NUMERIC;
        }
        if (st & PPG_ST_FEAS) {
            int screen = 0;
            if (P.is_qp && P.use_gram) {
                screen = kkt_cheb_gram(tw, act, &rad);
                if (screen < 0) { st |= PPG_ST_NUMERIC; screen = 0; }
                if (!screen && P.t > 1 && rad >= -PPG_RADIUS_BAND && rad < PPG_RADIUS_SCREEN) st |= PPG_ST_THIN;
            } else if (P.is_qp) {
                screen = 1;
            } else {
                screen = (P.ne + (int)act.size() == P.n) ? 1 : 0;  // mplp_program.py:472-473
            }
            if (screen) {
                st |= PPG_ST_OPT;
                double info[4];
                int r = emit_region(tw, act, laws.data(), rows.data(), flags.data(), info);
                if (r < 0) st |= PPG_ST_NUMERIC;
                if (r > 0) st |= PPG_ST_REGION;
                rad = info[1];
                if (r >= 0 && P.t > 1 && rad >= -PPG_RADIUS_BAND && rad <= PPG_RADIUS + PPG_RADIUS_BAND) st |= PPG_ST_THIN;
            }
        }
        (void)final_level_lp;
        status[ci] = st;
        if (aux) { aux[ci * 3] = ratio; aux[ci * 3 + 1] = margin; aux[ci * 3 + 2] = rad; }
    }
}

// feasibility LP of one active set with an optional replacement rhs column (the LP seen from another origin, as K2 starts
// from K2a's last iterate); returns the feasible flag, writes the pivot count.  twin_rows / twin_nfree / twin_t0 export the
// reduced feasibility rows for CPU emulations of the relaxation (tests/test_twin_k2a.py, step-count studies)
int twin_feas_rhs(void* h, const int* act, int k, const double* rhs, int* pivots_out) {
    Twin& tw = *(Twin*)h;
    const ReducedProgram& P = tw.P;
    Lp lp;
    lp_init(lp, P.R0, P.nfree);
    const int dc = P.nfree + 2;
    for (int i = 0; i < P.R0; ++i)
        for (int c = 0; c < dc; ++c) lp.at(i, c) = P.T0[(size_t)i * dc + c];
    if (rhs) for (int i = 0; i < P.R0; ++i) lp.at(i, 0) = rhs[i];
    for (int j = 0; j < k; ++j) { lp.rowflag[act[j]] = 2; lp.at(act[j], lp.js) = 0.0; }
    LpResult res = lp_maxmin(lp, -PPG_FEAS_TOL, false);
    if (pivots_out) *pivots_out = lp.pivots;
    if (res.code == PPG_LP_EARLY || res.code == PPG_LP_UNBOUNDED) return 1;
    if (res.code == PPG_LP_OPTIMAL) return res.beta >= -PPG_FEAS_TOL;
    return 0;
}
// K2w start vertex (host_math.hpp::build_walk_dictionary): returns wk_ok; D0 (nb x ld), bvar (nb), nvar (nfree) when ok
int twin_walk_dict(void* h, int* nb, int* ld, double* D0, int* bvar, int* nvar) {
    const ReducedProgram& P = ((Twin*)h)->P;
    *nb = P.wk_nb; *ld = P.wk_ld;
    if (!P.wk_ok) return 0;
    if (D0) std::memcpy(D0, P.wk_D0.data(), P.wk_D0.size() * sizeof(double));
    if (bvar) std::memcpy(bvar, P.wk_bvar.data(), P.wk_bvar.size() * sizeof(int));
    if (nvar) std::memcpy(nvar, P.wk_nvar.data(), P.wk_nvar.size() * sizeof(int));
    return 1;
}
// K2w on one level (candidates in lexicographic order): certified flags, returns the number of pivots
long twin_k2w(void* h, const uint64_t* masks, long ncand, uint8_t* certified) {
    std::vector<char> c;
    const long piv = k2w_walk_level(((Twin*)h)->P, masks, ncand, c);
    for (long i = 0; i < ncand; ++i) certified[i] = (uint8_t)c[i];
    return piv;
}
// ... with the witnesses (ncand x slots x ceil(R0 / 64) words, zero where there is none); closed: optional flags of candidates
// that arrive certified (inherited)
long twin_k2w_witness(void* h, const uint64_t* masks, long ncand, uint8_t* certified, uint64_t* witness, int slots,
                      const uint8_t* closed) {
    std::vector<char> c;
    const long piv = k2w_walk_level(((Twin*)h)->P, masks, ncand, c, witness, slots < 1 ? 1 : slots, closed);
    for (long i = 0; i < ncand; ++i) certified[i] = (uint8_t)c[i];
    return piv;
}
int twin_rows(void* h) { return ((Twin*)h)->P.R0; }
int twin_nfree(void* h) { return ((Twin*)h)->P.nfree; }
void twin_t0(void* h, double* out) {
    const ReducedProgram& P = ((Twin*)h)->P;
    for (size_t i = 0; i < P.T0.size(); ++i) out[i] = P.T0[i];
}

// K2a restatement for n candidates: flags[i] = 1 certified feasible / 0 not; steps[i]; for uncertified candidates the exact
// residuals G z* - h of the last iterate (what the kernel hands to K2) go to resid[i * R0 ..], zeros otherwise
static void twin_k2a_any(void* h, const uint64_t* masks, long ncand, int max_iter, int max_iter2, int32_t* flags, int32_t* steps,
                         double* resid, bool prefix) {
    Twin& tw = *(Twin*)h;
    const ReducedProgram& P = tw.P;
    std::vector<int> act;
    for (long ci = 0; ci < ncand; ++ci) {
        active_list(P, masks + ci * P.W, act);
        K2aOut o = prefix ? k2p_certify(P, act, max_iter, max_iter2) : k2a_certify(P, act, max_iter, max_iter2);
        flags[ci] = o.certified;
        steps[ci] = o.steps;
        if (resid) for (int r = 0; r < P.R0; ++r) resid[(size_t)ci * P.R0 + r] = o.resid.empty() ? 0.0 : o.resid[r];
    }
}

void twin_k2a(void* h, const uint64_t* masks, long ncand, int max_iter, int max_iter2, int32_t* flags, int32_t* steps,
              double* resid) {
    twin_k2a_any(h, masks, ncand, max_iter, max_iter2, flags, steps, resid, false);
}
// the prefix form (k2p_prefix.cu), same outputs
void twin_k2p(void* h, const uint64_t* masks, long ncand, int max_iter, int max_iter2, int32_t* flags, int32_t* steps,
              double* resid) {
    twin_k2a_any(h, masks, ncand, max_iter, max_iter2, flags, steps, resid, true);
}

int twin_emit(void* h, const uint64_t* mask, double* laws_out, double* rows_out, int32_t* flags_out, double* info) {
    Twin& tw = *(Twin*)h;
    std::vector<int> act;
    active_list(tw.P, mask, act);
    return emit_region(tw, act, laws_out, rows_out, flags_out, info);
}

// debug variant: runs every redundancy LP to optimality and reports its margin s*
int twin_emit_margins(void* h, const uint64_t* mask, double* laws_out, double* rows_out, int32_t* flags_out,
                      double* info, double* margins_out) {
    Twin& tw = *(Twin*)h;
    std::vector<int> act;
    active_list(tw.P, mask, act);
    return emit_region(tw, act, laws_out, rows_out, flags_out, info, margins_out);
}

}  // extern "C"
