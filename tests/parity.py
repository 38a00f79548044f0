"""Comparison helpers shared by the parity tests (reference semantics: SURVEY.md 3.4 / 7.1)."""
import numpy

REL_TOL = 1e-8  # north_star: law / half-space matrices within 1e-8 relative after PPOPT's own row normalisation


def rel_err(a, b):
    a = numpy.asarray(a, dtype=float)
    b = numpy.asarray(b, dtype=float)
    if a.shape != b.shape:
        return numpy.inf
    if a.size == 0:
        return 0.0
    scale = max(1.0, float(numpy.max(numpy.abs(b))))
    return float(numpy.max(numpy.abs(a - b))) / scale


def golden_regions(g):
    out = []
    for i in range(int(g['n_regions'])):
        out.append({k: g[f'r{i}_{k}'] for k in ('active_set', 'A', 'b', 'C', 'd', 'E', 'f', 'omega_set', 'lambda_set',
                                                  'regular_pos', 'regular_idx')})
    return out


def rows_match_as_sets(E1, f1, E2, f2, tol=REL_TOL):
    """every row of [E1|f1] has a partner in [E2|f2] within tol and vice versa; returns the unmatched rows"""
    R1 = numpy.hstack([numpy.asarray(E1, float), numpy.asarray(f1, float).reshape(-1, 1)])
    R2 = numpy.hstack([numpy.asarray(E2, float), numpy.asarray(f2, float).reshape(-1, 1)])

    def unmatched(X, Y):
        bad = []
        for r in X:
            if Y.shape[0] == 0:
                bad.append(r)
                continue
            d = numpy.max(numpy.abs(Y - r) / numpy.maximum(1.0, numpy.abs(r)), axis=1)
            if d.min() > tol:
                bad.append(r)
        return bad
    return unmatched(R1, R2), unmatched(R2, R1)


def masks_to_lists(masks, n_eq):
    m = numpy.ascontiguousarray(masks).view(numpy.uint64)
    m = m.reshape(m.shape[0], -1)
    bits = numpy.unpackbits(m.view(numpy.uint8).reshape(m.shape[0], -1), axis=1, bitorder='little')
    eq = list(range(n_eq))
    return [eq + (numpy.nonzero(r)[0] + n_eq).tolist() for r in bits]
