// Decision thresholds of the combinatorial path. Shared by the CUDA kernels and the CPU checker.
//
// Literal reference constants (must not be tuned):
//   PPG_ZERO_ROW   all |a_ij| <= 1e-8  -> numerically zero row   (utils/constraint_utilities.py:469-470)
//   PPG_RADIUS     Chebyshev radius > 1e-8 -> full dimensional    (utils/mpqp_utils.py:343)
//   PPG_WIDTH_1D   min + 1e-8 <= max                              (utils/mpqp_utils.py:320)
// LP-backend behaviour that the reference inherits from GLPK/HiGHS (primal feasibility tolerance):
//   PPG_FEAS_TOL   a row may be violated by at most 1e-7
// Engine-internal numerics (pivoting / anti-cycling), no reference counterpart:
//   PPG_PIV_TOL, PPG_OPT_TOL, PPG_HARRIS, PPG_DEGEN_STEP, PPG_BLAND_AFTER, PPG_RANK_TOL
#pragma once

#define PPG_ZERO_ROW 1e-8
#define PPG_RADIUS 1e-8
#define PPG_WIDTH_1D 1e-8
#define PPG_FEAS_TOL 1e-7
// redundancy LPs of gen_cr_from_active_set (mpqp_utils.py:143-178): a row is kept iff its hyperplane touches the
// region.  Calibrated on the runnable reference (HiGHS with presolve): it drops rows whose best margin is
// <= -2.7e-8 and keeps rows down to -4.8e-14 on every fixture except the stacked control-allocation family,
// where its own decisions overlap in [-2.3e-8, 5e-11] (DESIGN.md "weakly redundant rows").
#define PPG_REDUND_TOL 1e-9
// the Gram/Cholesky screen (K3/K4) passes candidates with half the radius to the LU-based final test
#define PPG_RADIUS_SCREEN 0.5e-8
// The reference decides "full dimensional" by comparing a radius that ITS LP BACKEND computed with 1e-8
// (mpqp_utils.py:343), and that backend only guarantees its answer to its own feasibility tolerance (1e-7): on
// ctrl_alloc_n5 level 4 HiGHS reports 2.78e-8 for 26 polytopes whose exact radius is -7.78e-9 (50-digit
// arithmetic).  The engine decides with the accurate radius and raises PPG_ST_THIN on every
// candidate whose radius lies within this band of the threshold, so that the discrepancy is never silent.
#define PPG_RADIUS_BAND 1e-7

#define PPG_PIV_TOL 1e-9
#define PPG_OPT_TOL 1e-12  // reduced-cost threshold: the objective error is tol x distance travelled (theta ranges of 1e2), so 1e-9 was too loose
#define PPG_HARRIS 1e-9
#define PPG_TINY 1e-12   // entries below this are treated as structural zeros in the ratio test
#define PPG_DEGEN_STEP 1e-12
#define PPG_BLAND_AFTER 8
// column-pivoted QR: rank deficient iff |R_kk| <= PPG_RANK_TOL * |R_11|
// (numpy.linalg.matrix_rank default: sigma_min <= sigma_max * max(k,n) * eps, constraint_utilities.py:236)
#define PPG_RANK_TOL 1e-11
#define PPG_RANK_BORDER_LO 1e-15   // below: exactly dependent rows (3e-16 on the MPC programs); above: re-decided by singular values
#define PPG_RANK_BORDER_HI 1e-7

// per-candidate status byte (bits 0-3 match tests/golden level*_status)
#define PPG_ST_RANK 1u      // LICQ holds: rank(A_active) == |active|
#define PPG_ST_FEAS 2u      // feasibility LP has a solution
#define PPG_ST_OPT 4u       // passed the optimality screen (theta-space polytope non-empty, full-dimensional)
#define PPG_ST_REGION 8u    // a critical region was emitted
#define PPG_ST_BORDER 16u   // a decision fell inside a borderline band (reported, never silently ignored)
#define PPG_ST_NUMERIC 32u  // iteration limit / singular KKT / non-finite value
#define PPG_ST_THIN 64u     // full-dimension decision taken inside PPG_RADIUS_BAND of the 1e-8 threshold (reported)
#define PPG_ST_PRE 128u       // transient: passed the K3 thread-per-candidate prefilter (cleared by k34_kernel)

// witness slots per candidate (ppgpu_level_eval_w): [0] the vertex that certified it (walk or inheritance), [1..] later
// vertices of the walk that hold it as well, written in turn.  Measured on the bench program (share of level 5 that inherits /
// ms per step): 1 slot 73 % / 240.7, 2 slots 85.5 % / 223.9, 3 slots 87.3 % / 220.3 - two it is (16 B per slot and candidate).  Also measured and rejected: slot 1 as the UNION of all
// later vertices (a "cover": certifies as well, but names no vertex a child could pass on - 82 % / 239.8; as a third slot
// next to two vertices 85.6 % / 230.5)
#define PPG_WITNESS_SLOTS 2

// LP return codes
#define PPG_LP_OPTIMAL 0
#define PPG_LP_EARLY 1
#define PPG_LP_UNBOUNDED 2
#define PPG_LP_INFEAS_EQ 3
#define PPG_LP_ITERLIM 4
