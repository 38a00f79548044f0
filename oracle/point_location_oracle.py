"""CPU restatement of the reference's point location and law evaluation.  TEST INFRASTRUCTURE - never imported by the product.

Follows, function by function:
  locate_upop      /root/reference/src/ppopt/upop/point_location.py:43-62   (first region with all(E theta <= f))
  locate_solution  /root/reference/src/ppopt/solution.py:75-88 + critical_region.py:81-84   (all(E theta - f < tol))
  evaluate         /root/reference/src/ppopt/critical_region.py:62-66   (A theta + b)
Pinned against outputs of the reference itself: tests/golden/pointloc/*.npz (oracle/gen_pointloc_golden.py).
"""
import numpy


def load_regions(g):
    """[(A, b, E, f)] from a golden archive (point-location or enumeration golden: same key names)"""
    return [(g[f'r{i}_A'], g[f'r{i}_b'], numpy.asarray(g[f'r{i}_E'], dtype=float), g[f'r{i}_f']) for i in range(int(g['n_regions']))]


def locate_upop(regions, theta):
    theta = numpy.asarray(theta, dtype=float).reshape(-1, 1)
    for j, (_, _, E, f) in enumerate(regions):
        if numpy.all(E @ theta <= f):
            return j
    return -1


def locate_solution(regions, theta, tol=1e-5):
    theta = numpy.asarray(theta, dtype=float).reshape(-1, 1)
    for j, (_, _, E, f) in enumerate(regions):
        if numpy.all(E @ theta - f < tol):
            return j
    return -1


def evaluate(regions, j, theta):
    if j < 0:
        return None
    A, b = regions[j][0], regions[j][1]
    return A @ numpy.asarray(theta, dtype=float).reshape(-1, 1) + b


def objective(prog, x, theta):
    """MPQP_Program.evaluate_objective (/root/reference/src/ppopt/mpqp_program.py:44-57) / MPLP_Program's (mplp_program.py:158-160);
    prog: dict with Q (or None), H, c, and optionally c_c, c_t, Q_t (zero by default, mplp_program.py:71-81)"""
    theta = numpy.asarray(theta, dtype=float).reshape(-1, 1)
    t = theta.shape[0]
    c_c = prog.get('c_c', numpy.zeros((1, 1)))
    c_t = prog.get('c_t', numpy.zeros((t, 1)))
    Q_t = prog.get('Q_t', numpy.zeros((t, t)))
    v = theta.T @ prog['H'].T @ x + prog['c'].T @ x + c_c + c_t.T @ theta + 0.5 * theta.T @ Q_t @ theta
    if prog.get('Q') is not None:
        v = v + 0.5 * x.T @ prog['Q'] @ x
    return float(v[0, 0])


def locate_overlap(regions, prog, theta, tol=None, return_gap=False):
    """Solution.get_region_overlap (solution.py:90-112; tol = point_location_tolerance) or, with tol=None,
    upop.PointLocation in overlapping mode (upop/point_location.py:43-52,68-84): lowest objective, ties to the later region"""
    theta = numpy.asarray(theta, dtype=float).reshape(-1, 1)
    best, best_obj, objs = -1, float('inf'), []
    for j, (A, b, E, f) in enumerate(regions):
        inside = numpy.all(E @ theta <= f) if tol is None else numpy.all(E @ theta - f < tol)
        if inside:
            obj = objective(prog, A @ theta + b, theta)
            objs.append(obj)
            if obj <= best_obj:
                best, best_obj = j, obj
    if return_gap:
        objs = sorted(objs)
        return best, (objs[1] - objs[0]) / max(1.0, abs(objs[0])) if len(objs) > 1 else numpy.inf
    return best
