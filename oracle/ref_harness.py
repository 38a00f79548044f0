"""Imports the UNMODIFIED reference (PPOPT) from /root/reference/src under the LP shim.

TEST INFRASTRUCTURE.  Only usable in the build container (where /root/reference exists);
used by oracle/gen_golden.py to produce tests/golden/* and by local (non-gpu) tests that
skip when the reference is absent.  Nothing under ppopt_b200/ imports this.
"""
import os
import sys

REF_SRC = '/root/reference/src'
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_shim')


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, 'ppopt'))


def load():
    """Returns the reference `ppopt` package (shimmed).  Idempotent."""
    if not available():
        raise RuntimeError('reference not present at ' + REF_SRC)
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    if REF_SRC not in sys.path:
        sys.path.insert(1, REF_SRC)
    import cvxopt  # noqa: F401  (the shim; must be in sys.modules before ppopt.solver is imported)
    import ppopt
    import ppopt.solver as _solver
    import ppopt.solver_interface.solver_interface as _si
    from ppopt.solver_interface.cvxopt_interface import solve_lp_cvxopt
    # CriticalRegion.is_full_dimension() -> chebyshev_ball(..., deterministic_solver='gurobi')
    # (critical_region.py:97, chebyshev_ball.py:11); gurobi is absent, so point that name at the shim.
    _si.solve_lp_gurobi = solve_lp_cvxopt
    _solver.solve_lp_gurobi = solve_lp_cvxopt
    return ppopt
