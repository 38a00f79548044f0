"""Golden vectors for batched point location, produced by the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the build container only:  python oracle/gen_pointloc_golden.py
For each program: the reference's combinatorial solve (under the LP shim), then for a seeded cloud of theta points
  * upop.PointLocation(solution).locate(theta)      (/root/reference/src/ppopt/upop/point_location.py:43-62,  E theta <= f)
  * Solution.get_region_no_overlap(theta)           (/root/reference/src/ppopt/solution.py:75-88, is_inside with tol 1e-5,
                                                     critical_region.py:81-84)
  * region.evaluate(theta)                          (critical_region.py:62-74)
are recorded next to the region matrices in tests/golden/pointloc/<name>.npz.
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness  # noqa: E402

ppopt = ref_harness.load()
from ppopt.mp_solvers import mpqp_combinatorial  # noqa: E402
from ppopt.upop.point_location import PointLocation  # noqa: E402

from gen_golden import build_reference_program  # noqa: E402
import problems  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
NAMES = ['factory_mpqp', 'rand_6_3_12_s1', 'mpc_n3', 'ctrl_alloc_n1']


def theta_box(prog):
    """bounding box of Theta from the single-variable rows of A_t theta <= b_t (all fixtures use box Thetas)"""
    t = prog.num_t()
    lo, hi = numpy.full(t, -numpy.inf), numpy.full(t, numpy.inf)
    for row, rhs in zip(prog.A_t, prog.b_t.ravel()):
        nz = numpy.nonzero(row)[0]
        if len(nz) == 1:
            j = nz[0]
            if row[j] > 0:
                hi[j] = min(hi[j], rhs / row[j])
            else:
                lo[j] = max(lo[j], rhs / row[j])
    lo = numpy.where(numpy.isfinite(lo), lo, numpy.where(numpy.isfinite(hi), hi - 20.0, -10.0))
    hi = numpy.where(numpy.isfinite(hi), hi, lo + 20.0)
    return lo, hi


def generate(name, n_points=600, seed=7):
    prog = build_reference_program(problems.CONFIGS[name]())
    sol = mpqp_combinatorial.solve(prog)
    sol.is_overlapping = False   # positive definite mpQPs: what solve_mpqp leaves (solve_mpqp.py:103-112)
    regions = sol.critical_regions
    lo, hi = theta_box(prog)
    rng = numpy.random.default_rng(seed)
    pad = 0.15 * (hi - lo)
    thetas = rng.uniform(lo - pad, hi + pad, size=(n_points, prog.num_t()))   # ~1/3 of the cloud falls outside Theta
    # plus points 3e-6 OUTSIDE the nearest facet of their region: inside for Solution.get_region (tol 1e-5), outside that
    # region for upop.PointLocation (E theta <= f) - pins the two acceptance rules separately
    extra = []
    for th in thetas[:400]:
        r = sol.get_region_no_overlap(th.reshape(-1, 1))
        if r is None or len(extra) >= 120:
            continue
        E, f = numpy.asarray(r.E, dtype=float), numpy.asarray(r.f, dtype=float).ravel()
        slack = f - E @ th
        i = int(numpy.argmin(slack))
        extra.append(th + (slack[i] + 3e-6) * E[i] / float(E[i] @ E[i]))
    thetas = numpy.vstack([thetas, numpy.array(extra).reshape(-1, prog.num_t())])
    n_points = thetas.shape[0]
    pl = PointLocation(sol)
    idx_upop = numpy.array([int(pl.locate(th.reshape(-1, 1))) for th in thetas], dtype=numpy.int32)
    idx_sol = numpy.full(n_points, -1, dtype=numpy.int32)
    x_sol = numpy.full((n_points, prog.num_x()), numpy.nan)
    for p, th in enumerate(thetas):
        r = sol.get_region_no_overlap(th.reshape(-1, 1))
        if r is not None:
            idx_sol[p] = next(i for i, q in enumerate(regions) if q is r)
            x_sol[p] = r.evaluate(th.reshape(-1, 1)).ravel()
    # the overlapping rule (what solve_mpqp leaves on every solution, solve_mpqp.py:105-112: MPQP_Program IS an
    # MPLP_Program): among the containing regions the one with the lowest objective, ties to the later region
    # (solution.py:90-112, upop/point_location.py:43-52,68-84)
    sol.is_overlapping = True
    plo = PointLocation(sol)
    idx_upop_ov = numpy.array([int(plo.locate(th.reshape(-1, 1))) for th in thetas], dtype=numpy.int32)
    idx_sol_ov = numpy.full(n_points, -1, dtype=numpy.int32)
    for p, th in enumerate(thetas):
        r = sol.get_region_overlap(th.reshape(-1, 1))
        if r is not None:
            idx_sol_ov[p] = next(i for i, q in enumerate(regions) if q is r)
    out = dict(thetas=thetas, idx_upop=idx_upop, idx_solution=idx_sol, x_solution=x_sol, n_regions=numpy.int64(len(regions)),
               idx_upop_overlap=idx_upop_ov, idx_solution_overlap=idx_sol_ov,
               point_location_tolerance=numpy.float64(sol.point_location_tolerance))
    for i, r in enumerate(regions):
        out[f'r{i}_A'], out[f'r{i}_b'], out[f'r{i}_E'], out[f'r{i}_f'] = r.A, r.b, numpy.asarray(r.E, dtype=float), r.f
    os.makedirs(os.path.join(OUT, 'pointloc'), exist_ok=True)
    numpy.savez_compressed(os.path.join(OUT, 'pointloc', f'{name}.npz'), **out)
    print(f'[{name}] {len(regions)} regions, {n_points} points: located {int((idx_upop >= 0).sum())} (upop), '
          f'{int((idx_sol >= 0).sum())} (Solution, tol 1e-5), differ on {int((idx_upop != idx_sol).sum())}; overlap rule changes '
          f'{int((idx_sol_ov != idx_sol).sum())} (Solution) / {int((idx_upop_ov != idx_upop).sum())} (upop)', flush=True)


if __name__ == '__main__':
    for nm in sys.argv[1:] or NAMES:
        generate(nm)
