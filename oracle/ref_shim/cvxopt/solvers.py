"""`cvxopt.solvers` stand-in backed by scipy.optimize.linprog(method='highs')."""
import numpy
from scipy.optimize import linprog

options = {}


def lp(c, G=None, h=None, A=None, b=None, solver=None, options=None, **kw):
    c = numpy.asarray(c, dtype=float).ravel()
    n = c.size
    G_ = None if G is None else numpy.asarray(G, dtype=float).reshape(-1, n)
    h_ = None if h is None else numpy.asarray(h, dtype=float).ravel()
    A_ = None if A is None else numpy.asarray(A, dtype=float).reshape(-1, n)
    b_ = None if b is None else numpy.asarray(b, dtype=float).ravel()
    res = linprog(c, A_ub=G_, b_ub=h_, A_eq=A_, b_eq=b_, bounds=(None, None), method='highs')
    if res.status == 0:
        status = 'optimal'
    elif res.status == 2:
        status = 'primal infeasible'
    elif res.status == 3:
        status = 'dual infeasible'
    else:
        status = 'unknown'
    out = {'status': status, 'x': None, 's': None, 'z': None, 'y': None, 'primal objective': None}
    if status == 'optimal':
        out['x'] = res.x.reshape(-1, 1)
        out['primal objective'] = float(res.fun)
        if G_ is not None:
            out['s'] = (h_ - G_ @ res.x).reshape(-1, 1)
            out['z'] = -numpy.asarray(res.ineqlin.marginals).reshape(-1, 1)
        else:
            out['s'] = numpy.zeros((0, 1))
            out['z'] = numpy.zeros((0, 1))
        if A_ is not None:
            out['y'] = -numpy.asarray(res.eqlin.marginals).reshape(-1, 1)
        else:
            out['y'] = numpy.zeros((0, 1))
    return out


def qp(*a, **k):
    raise NotImplementedError("the combinatorial path never calls a QP solver")
