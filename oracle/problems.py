"""Raw problem data for the benchmark / parity configurations (BASELINE.json `configs`).  TEST INFRASTRUCTURE: fixture
definitions for oracle/gen_*golden.py and tests/; the product package does not import this.

Each builder returns a plain dict of numpy arrays in the reference's constructor
convention  ``min 1/2 x'Qx + theta'H'x + c'x  s.t.  Ax <= b + F theta,  A_t theta <= b_t``
(/root/reference/src/ppopt/mpqp_program.py:16-27) plus ``equality_indices`` and ``kind``
('qp' | 'lp').  The dicts feed either the reference's MPQP_Program/MPLP_Program (golden
generation, oracle/gen_golden.py) or this package's own program classes.

Sources of the data (definitions only - nothing here is computed by the reference):
  factory_mpqp        /root/reference/tests/test_fixtures.py:17-31
  simple_mpqp_1d      /root/reference/tests/test_fixtures.py:51-63
  simple_mplp         /root/reference/tests/test_fixtures.py:191-203
  portfolio_analog    /root/reference/tests/test_fixtures.py:262-281
  doc_portfolio       /root/reference/doc/portfolio.rst:24-45
  transport_mplp      /root/reference/doc/mplp_tut.rst:28-36
  mpc_double_integrator  /root/reference/doc/mpc.rst:69-116 generalised to horizon N
  control_allocation  /root/reference/doc/control_allocation_example.rst:21-228 (4 rotors,
                      N stacked steps with a rate penalty)
  random_mpqp         /root/reference/src/ppopt/problem_generator.py:25-78 (same draws, same order)
"""
from typing import Dict, Optional

import numpy


def _col(x):
    return numpy.asarray(x, dtype=float).reshape(-1, 1)


def factory_mpqp() -> Dict:
    A = numpy.array([[1, 1, 0, 0], [0, 0, 1, 1], [-1, 0, -1, 0], [0, -1, 0, -1], [-1, 0, 0, 0], [0, -1, 0, 0],
                     [0, 0, -1, 0], [0, 0, 0, -1]], dtype=float)
    b = _col([350, 600, 0, 0, 0, 0, 0, 0])
    c = 25.0 * _col([1, 1, 1, 1])
    F = numpy.array([[0, 0], [0, 0], [-1, 0], [0, -1], [0, 0], [0, 0], [0, 0], [0, 0]], dtype=float)
    Q = 2.0 * numpy.diag([153.0, 162.0, 162.0, 126.0])
    A_t = numpy.vstack((numpy.eye(2), -numpy.eye(2)))
    b_t = _col([1000, 1000, 0, 0])
    H = numpy.zeros((4, 2))
    return dict(kind='qp', A=A, b=b, c=c, H=H, Q=Q, A_t=A_t, b_t=b_t, F=F, equality_indices=[])


def transport_mplp() -> Dict:
    d = factory_mpqp()
    d.pop('Q')
    d['kind'] = 'lp'
    d['c'] = _col([178, 187, 187, 151])
    return d


def simple_mpqp_1d() -> Dict:
    return dict(kind='qp', Q=numpy.array([[1.0]]), A=numpy.array([[1.0], [-1.0]]), b=_col([5, 0]), c=_col([0]),
                F=numpy.array([[1.0], [1.0]]), A_t=numpy.array([[-1.0], [1.0]]), b_t=_col([0, 1]),
                H=numpy.zeros((1, 1)), equality_indices=[])


def simple_mplp() -> Dict:
    A = numpy.array([[0, 1, 1], [1, 0, 0], [-1, 0, 0], [1, -1, 0], [1, 0, -1]], dtype=float)
    return dict(kind='lp', A=A, b=_col([1, 0, 0, 0, 0]), F=_col([0, 1, 0, 0, 0]), c=_col([-3, 0, 0]),
                H=numpy.zeros((3, 1)), A_t=_col([1, 1]), b_t=_col([2, 2]), equality_indices=[])


def portfolio_analog() -> Dict:
    n = 8
    mu = [0.09551451, 0.00317183, 0.06799116, 0.12334409, 0.10235298, 0.0754139, 0.00730871, 0.11324299]
    A = numpy.vstack([numpy.ones((1, n)), numpy.array(mu).reshape(1, n), -numpy.eye(n)])
    b = _col([1, 0] + [0] * n)
    F = numpy.vstack([[0.0], [1.0], numpy.zeros((n, 1))])
    return dict(kind='qp', A=A, b=b, F=F, A_t=numpy.array([[-1.0], [1.0]]), b_t=_col([-min(mu), max(mu)]),
                Q=numpy.diag([float(i + 1) for i in range(n)]), c=numpy.zeros((n, 1)), H=numpy.zeros((n, 1)),
                equality_indices=[0, 1], post_process=False)


def doc_portfolio(num_assets: int = 10, seed: int = 123456789) -> Dict:
    rs = numpy.random.RandomState(seed)  # same legacy stream as numpy.random.seed(seed)
    S = rs.randn(num_assets, num_assets)
    S = S @ S.T / 10
    mu = rs.rand(num_assets)
    A = numpy.vstack([numpy.ones((1, num_assets)), mu.reshape(1, -1), -numpy.eye(num_assets)])
    b = _col([1, 0] + [0] * num_assets)
    F = numpy.vstack([[0.0], [1.0], numpy.zeros((num_assets, 1))])
    return dict(kind='qp', A=A, b=b, F=F, A_t=numpy.array([[-1.0], [1.0]]), b_t=_col([-mu.min(), mu.max()]), Q=S,
                c=numpy.zeros((num_assets, 1)), H=numpy.zeros((num_assets, 1)), equality_indices=[0, 1])


def mpc_double_integrator(N: int = 3) -> Dict:
    """x_{k+1} = A_ss x_k + B_ss u_k, box bounds |x|<=4, |u|<=1, theta = x_0, decision [x_1..x_N, u_0..u_{N-1}]."""
    A_ss = numpy.array([[1.0, 1.0], [0.0, 1.0]])
    B_ss = numpy.array([[0.5], [1.0]])
    nx, nu = 2 * N, N
    n = nx + nu
    A_eq = numpy.zeros((2 * N, n))
    F_eq = numpy.zeros((2 * N, 2))
    for k in range(N):
        r = slice(2 * k, 2 * k + 2)
        A_eq[r, 2 * k:2 * k + 2] = numpy.eye(2)
        if k > 0:
            A_eq[r, 2 * (k - 1):2 * k] = -A_ss
        A_eq[r, nx + k:nx + k + 1] = -B_ss
    F_eq[0:2] = A_ss
    A_in = numpy.vstack([numpy.eye(n), -numpy.eye(n)])
    ub = numpy.concatenate([4.0 * numpy.ones(nx), numpy.ones(nu)])
    b_in = numpy.concatenate([ub, ub]).reshape(-1, 1)
    return dict(kind='qp', A=numpy.vstack([A_eq, A_in]), b=numpy.vstack([numpy.zeros((2 * N, 1)), b_in]),
                F=numpy.vstack([F_eq, numpy.zeros((2 * n, 2))]), A_t=numpy.vstack([numpy.eye(2), -numpy.eye(2)]),
                b_t=4.0 * numpy.ones((4, 1)), Q=numpy.eye(n), c=numpy.zeros((n, 1)), H=numpy.zeros((n, 2)),
                equality_indices=list(range(2 * N)))


def control_allocation(steps: int = 1, rho: float = 1.0) -> Dict:
    """4-rotor control allocation; `steps` stacked allocations of the same command coupled by rho*||u_k-u_{k-1}||^2."""
    n, m = 4, 4
    g, r = 9.8, 0.35
    rotDir = numpy.array([+1.0, -1.0, +1.0, -1.0])
    phi = numpy.linspace(0.0, 2.0 * numpy.pi, n + 1)[:-1]
    xR, yR = numpy.sin(phi), numpy.cos(phi)
    Ct, FoM = 0.014, 0.7
    Cq = Ct ** 1.5 / FoM / numpy.sqrt(2.0)
    J = numpy.zeros((m, n))
    dT = Ct / (r * Cq)
    J[0, :] = -dT
    J[1, :] = -dT * yR
    J[2, :] = dT * xR
    J[3, :] = rotDir
    FMTrim = numpy.array([-m * g, 0.0, 0.0, 0.0])  # the doc rebinds m=4 before this line (rst:24,30,79)
    xTrim = (numpy.linalg.pinv(J) @ FMTrim.reshape(4, 1)).reshape(n)
    W = numpy.diag([20.0, 100.0, 100.0, 5.0])
    Q1 = J.T @ W @ J + J.T @ J
    c1 = -J.T @ J @ xTrim.reshape(n, 1)
    H1 = -J.T @ W
    xMax = 1.4 * numpy.mean(xTrim) * numpy.ones(n)
    lo = numpy.array([-1.2 * m * g, -15.0, -15.0, -3.0]) * 1.1
    hi = numpy.array([-0.8 * m * g, 15.0, 15.0, 3.0]) * 1.1
    A_t = numpy.vstack([-numpy.eye(m), numpy.eye(m)])
    b_t = numpy.concatenate([-lo, hi]).reshape(-1, 1)
    N = steps
    nn = n * N
    Q = numpy.zeros((nn, nn))
    c = numpy.zeros((nn, 1))
    H = numpy.zeros((nn, m))
    for k in range(N):
        s = slice(n * k, n * k + n)
        Q[s, s] += Q1
        c[s] += c1
        H[s] = H1
        if N > 1:
            # rho * ||u_k - u_{k-1}||^2 with u_{-1} = xTrim (objective is 1/2 x'Qx + c'x, so factor 2)
            Q[s, s] += 2.0 * rho * numpy.eye(n)
            if k == 0:
                c[s] += -2.0 * rho * xTrim.reshape(n, 1)
            else:
                p = slice(n * (k - 1), n * k)
                Q[p, p] += 2.0 * rho * numpy.eye(n)
                Q[s, p] += -2.0 * rho * numpy.eye(n)
                Q[p, s] += -2.0 * rho * numpy.eye(n)
    A = numpy.vstack([-numpy.eye(nn), numpy.eye(nn)])
    b = numpy.concatenate([numpy.zeros(nn), numpy.tile(xMax, N)]).reshape(-1, 1)
    return dict(kind='qp', A=A, b=b, F=numpy.zeros((2 * nn, m)), A_t=A_t, b_t=b_t, Q=Q, c=c, H=H,
                equality_indices=[])


def random_mpqp(x: int = 2, t: int = 2, m: int = 10, seed: Optional[int] = None, kind: str = 'qp') -> Dict:
    """Same random draws, in the same order, as the reference generator (problem_generator.py:25-78)."""
    prng = numpy.random.default_rng(seed)
    Q = prng.random((x, x))
    Q = Q.T @ Q + numpy.eye(x)
    draw = lambda: prng.random(1)
    range_value = numpy.round(20 * draw() + 5)
    x_border = numpy.round(8 * draw() + 1) / 10
    x_shift = numpy.round(8 * draw() + 1) / 10
    t_border = numpy.round(8 * draw() + 1) / 10
    t_shift = numpy.round(8 * draw() + 1) / 10
    c = (prng.random((x, 1)) - .5) / draw()
    ev = numpy.linalg.eigvals(Q)
    rng_ = range_value * (max(ev) - min(ev))
    A = numpy.zeros((m, x))
    F = numpy.zeros((m, t))
    for i in range(m):
        ok = False
        while not ok:
            idx = prng.random(x) >= x_border
            A[i][idx] = numpy.floor((prng.random(sum(idx)) - x_shift) * rng_)
            ok = bool(numpy.any(A[i] != 0))
        idx = prng.random(t) >= t_border
        F[i][idx] = numpy.floor((prng.random(sum(idx)) - t_shift) * rng_)
    A = numpy.vstack([A, numpy.eye(x), -numpy.eye(x)])
    F = numpy.vstack([F, numpy.zeros((2 * x, t))])
    b = numpy.vstack([prng.random((m, 1)) / prng.random(1), 1e7 * numpy.ones((2 * x, 1))])
    A_t = numpy.vstack([numpy.eye(t), -numpy.eye(t)])
    b_t = rng_ * numpy.ones((2 * t, 1))
    out = dict(kind=kind, A=A, b=b, c=c, H=numpy.zeros((x, t)), Q=Q, A_t=A_t, b_t=b_t, F=F, equality_indices=[])
    if kind == 'lp':
        out.pop('Q')
    return out


CONFIGS = {
    'factory_mpqp': factory_mpqp,
    'transport_mplp': transport_mplp,
    'simple_mpqp_1d': simple_mpqp_1d,
    'simple_mplp': simple_mplp,
    'portfolio_analog': portfolio_analog,
    'doc_portfolio': doc_portfolio,
    'mpc_n3': lambda: mpc_double_integrator(3),
    'mpc_n5': lambda: mpc_double_integrator(5),
    'mpc_n7': lambda: mpc_double_integrator(7),
    'mpc_n10': lambda: mpc_double_integrator(10),
    'ctrl_alloc_n1': lambda: control_allocation(1),
    'ctrl_alloc_n2': lambda: control_allocation(2),
    'ctrl_alloc_n5': lambda: control_allocation(5),
    'rand_6_3_12_s1': lambda: random_mpqp(6, 3, 12, 1),
    'rand_5_3_10_s2': lambda: random_mpqp(5, 3, 10, 2),
    'rand_lp_4_2_8_s3': lambda: random_mpqp(4, 2, 8, 3, kind='lp'),
    'synthetic_30_6_40_s0': lambda: random_mpqp(30, 6, 40, 0),
    # exercises the wide template instantiations: n' > 32, t > 6, more than 128 rows
    'rand_wide_40_8_90_s5': lambda: random_mpqp(40, 8, 90, 5),
}
