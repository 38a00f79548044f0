"""Stand-in for the `cvxopt` package, used ONLY to run the unmodified reference here.

TEST INFRASTRUCTURE - never imported by the product path (ppopt_b200/).

The reference's LP seam calls ``cvxopt.solvers.lp(..., solver='glpk')``
(/root/reference/src/ppopt/solver_interface/cvxopt_interface.py:205-206) and reads the
keys ``status, x, s, z, y, 'primal objective'`` (same file :18-51).  cvxopt/GLPK are not
installed in this image, so this shim answers the same call with scipy's HiGHS.  Placing
this directory first on ``sys.path`` also makes ``'cvxopt' in sys.modules`` true, which is
how the reference chooses ``glpk`` as its default LP backend (solver.py:19-73).
"""
import numpy

from . import solvers  # noqa: F401


def matrix(x, *args, **kwargs):
    return numpy.array(x, dtype=float)
