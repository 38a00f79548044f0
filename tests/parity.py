"""Comparison helpers shared by the parity tests (reference semantics: SURVEY.md 3.4 / 7.1)."""
import numpy

REL_TOL = 1e-8  # north_star: law / half-space matrices within 1e-8 relative after PPOPT's own row normalisation
ST_THIN = 64    # csrc/tolerances.h PPG_ST_THIN: full-dimension decision inside the LP backend's tolerance band


def check_status_bits(mine, ref, what=''):
    """rank (1) and feasible (2) must agree on every candidate.  The region bit (8) must agree too, except on candidates
    that the engine itself flagged PPG_ST_THIN: there the reference's answer is its LP backend's rounding (HiGHS reports a
    radius of 2.78e-8 for polytopes whose exact radius is -7.78e-9), the engine answers with the
    accurate radius and says so.  Returns the indices of such tolerated candidates."""
    mine = numpy.asarray(mine)
    ref = numpy.asarray(ref)
    for bit, name in ((1, 'rank'), (2, 'feasible')):
        bad = numpy.nonzero((mine & bit) != (ref & bit))[0]
        assert bad.size == 0, f'{what}: {name} differs at {bad[:5].tolist()}'
    bad = numpy.nonzero((mine & 8) != (ref & 8))[0]
    untolerated = [int(i) for i in bad if not (mine[i] & ST_THIN)]
    assert not untolerated, f'{what}: region decision differs at {untolerated[:5]} (not flagged thin)'
    return bad


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
