"""Imports the UNMODIFIED reference (PPOPT) under the LP shim.

TEST INFRASTRUCTURE.  The reference is taken from ``baseline/_ref`` (the offline ``pip install --target`` of
/root/reference recorded in DESIGN.md section 9; git-ignored, but it travels to the GPU box with the snapshot), else from
/root/reference/src (build container only).  Used by oracle/gen_golden.py to produce tests/golden/*, by bench.py's
reference arm, and by tests that skip when the reference is absent.  Nothing under ppopt_b200/ imports this.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.path.join(_ROOT, 'baseline', '_ref'), '/root/reference/src']
REF_SRC = next((p for p in _CANDIDATES if os.path.isdir(os.path.join(p, 'ppopt'))), _CANDIDATES[-1])
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
