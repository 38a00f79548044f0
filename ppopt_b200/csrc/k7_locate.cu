// K7 - batched point location and law evaluation (the step right after the enumeration).
//
// Replaces, for a whole batch of parameter points, the per-point Python loops of
//   Solution.get_region_no_overlap / evaluate   (/root/reference/src/ppopt/solution.py:44-88; region.is_inside with a
//                                                 tolerance: all(E theta - f < tol), critical_region.py:81-84)
//   upop.PointLocation.locate / evaluate        (/root/reference/src/ppopt/upop/point_location.py:43-62,121-133; direct
//                                                 enumeration over the stacked half-spaces with E theta <= f)
// Both return the FIRST region, in solution order, that contains the point; x*(theta) = A theta + b of that region
// (critical_region.py:62-66).  One warp per point: lanes take the rows of one region at a time, a vote decides, the
// scan stops at the first hit.  The stacked rows (a few hundred KB) stay in L2; the kernel is bound by fp64 FMA issue
// and L2 latency, not HBM (8 t B in, 4 + 8 n B out per point).
#include "common.cuh"
#include "launch.h"

namespace ppgpu {

constexpr int K7_MAXT = 32;
constexpr int K7_MAXN = 128;   // variables, overlapping rule only (x* of a candidate region is staged in shared memory)

// overlap == 0: first containing region.  overlap == 1: among the containing regions the one with the lowest objective
// 1/2 x'Qx + theta'H'x + c'x at x = A theta + b, ties to the later region (solution.py:90-112, point_location.py:68-84;
// the region-independent terms c_c + c_t'theta + 1/2 theta'Q_t theta of evaluate_objective are left out: they do not move
// the arg-min).  On a shared facet two laws give objectives that agree to rounding, so WHICH of the tied regions wins is
// noise in the reference as well; the value of the winner is what is comparable.
__global__ void __launch_bounds__(128)
locate_points_kernel(const double* __restrict__ theta, long long n_points, int t, const double* __restrict__ rows,
                     const long long* __restrict__ row_off, long long n_regions, const double* __restrict__ laws, int n_x,
                     int use_tol, double tol, int overlap, const double* __restrict__ Qm, const double* __restrict__ Hm,
                     const double* __restrict__ cv, int* __restrict__ region_out, double* __restrict__ x_out) {
    __shared__ double th_s[4][K7_MAXT];
    __shared__ double x_s[4][K7_MAXN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * 4;
    const int t1 = t + 1;
    double* th = th_s[warp];
    double* xs = x_s[warp];
    auto law_value = [&](long long r, int i) {
        const double* law = laws + ((size_t)r * n_x + i) * t1;   // [b_i | A_i1 .. A_it]
        double v = __ldg(law);
        for (int j = 0; j < t; ++j) v = fma(__ldg(law + 1 + j), th[j], v);
        return v;
    };
    for (long long p = (long long)blockIdx.x * 4 + warp; p < n_points; p += nw) {
        __syncwarp();
        if (lane < t) th[lane] = theta[p * t + lane];
        __syncwarp();
        long long found = -1;
        double best = __longlong_as_double(0x7ff0000000000000ll);   // +inf
        for (long long r = 0; r < n_regions; ++r) {
            const long long lo = row_off[r], hi = row_off[r + 1];
            bool bad = false;
            for (long long i = lo + lane; i < hi; i += 32) {
                const double* row = rows + i * t1;   // [f | a_1 .. a_t]
                double s = 0.0;
                for (int j = 0; j < t; ++j) s = fma(__ldg(row + 1 + j), th[j], s);
                const double f = __ldg(row);
                const bool ok = use_tol ? (s - f < tol) : (s <= f);   // NaN: not inside, as numpy.all(... < tol)
                bad = bad || !ok;
            }
            if (__any_sync(0xffffffffu, bad)) continue;
            if (!overlap) { found = r; break; }
            for (int i = lane; i < n_x; i += 32) xs[i] = law_value(r, i);
            __syncwarp();
            double part = 0.0;
            for (int i = lane; i < n_x; i += 32) {
                double gi = __ldg(cv + i);
                for (int k = 0; k < t; ++k) gi = fma(__ldg(Hm + (size_t)i * t + k), th[k], gi);
                if (Qm != nullptr) {
                    double q = 0.0;
                    for (int j = 0; j < n_x; ++j) q = fma(__ldg(Qm + (size_t)i * n_x + j), xs[j], q);
                    gi = fma(0.5, q, gi);
                }
                part = fma(xs[i], gi, part);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            __syncwarp();   // xs is rewritten for the next containing region
            if (part <= best) { best = part; found = r; }
        }
        if (lane == 0) region_out[p] = (int)found;
        if (x_out != nullptr) {
            for (int i = lane; i < n_x; i += 32)
                x_out[p * n_x + i] = found >= 0 ? law_value(found, i) : __longlong_as_double(0x7ff8000000000000ll);
        }
    }
}

cudaError_t launch_locate(const double* theta, long long n_points, int t, const double* rows, const long long* row_off,
                          long long n_regions, const double* laws, int n_x, int use_tol, double tol, int overlap,
                          const double* Qm, const double* Hm, const double* cv, int* region_out, double* x_out, int sm_count,
                          cudaStream_t st) {
    if (n_points <= 0) return cudaSuccess;
    if (t < 1 || t > K7_MAXT) return cudaErrorInvalidValue;
    if (overlap && (n_x < 1 || n_x > K7_MAXN || !Hm || !cv || !laws)) return cudaErrorInvalidValue;
    long long grid = (n_points + 3) / 4;
    const long long cap = (long long)sm_count * 16;
    if (grid > cap) grid = cap;
    locate_points_kernel<<<(unsigned)grid, 128, 0, st>>>(theta, n_points, t, rows, row_off, n_regions, laws, n_x, use_tol, tol,
                                                         overlap, Qm, Hm, cv, region_out, x_out);
    return cudaGetLastError();
}

}  // namespace ppgpu
