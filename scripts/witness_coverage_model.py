"""CPU model of witness inheritance on the bench program (synthetic 100 x 30 x 6), levels 3 -> 4 -> 5, with the CPU checker's
sequential walk (oracle/twin.cpp::k2w_walk_level, the restatement of csrc/k2w_walk.cu incl. the second witness slot):

    python scripts/witness_coverage_model.py [sample]

  level 3: walked in full (the golden level of the unmodified reference), witnesses kept;
  level 4: every candidate K6 would generate; those covered by a parent's witness arrive closed (and inherit that witness),
           the rest is walked; closed candidates are revisited (slot 1);
  level 5: a random sample of candidates; share covered by a parent's witness, with slot 0 alone and with both slots.

This is the model quoted in DESIGN.md section 3 (the device measures 68 % at level 4 and 73 % / 85.5 % at level 5).
TEST INFRASTRUCTURE (imports oracle/): never used by the product."""
import itertools
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from twin_binding import Twin  # noqa: E402

PATH = os.path.join(ROOT, 'tests', 'golden', 'synthetic_30_6_40_s0.npz')
M = 100


def keys(S):
    k = numpy.zeros(S.shape[0], dtype=numpy.int64)
    for i in range(S.shape[1]):
        k = k * M + S[:, i].astype(numpy.int64)
    return k


def to_masks(S):
    m = numpy.zeros((S.shape[0], 2), dtype=numpy.uint64)
    rows = numpy.arange(S.shape[0])
    for i in range(S.shape[1]):
        r = S[:, i].astype(numpy.uint64)
        m[rows, (r >> numpy.uint64(6)).astype(numpy.int64)] |= numpy.uint64(1) << (r & numpy.uint64(63))
    return m


def children(S, feasible_keys):
    """next level in lexicographic order: parent + a larger row, every (k)-subset feasible (the rule of K6)"""
    k = S.shape[1]
    parts = []
    for last in range(M):
        sel = S[S[:, -1] < last]
        parts.append(numpy.concatenate([sel, numpy.full((sel.shape[0], 1), last, dtype=S.dtype)], axis=1))
    C = numpy.concatenate(parts)
    C = C[numpy.argsort(keys(C), kind='stable')]
    ok = numpy.ones(C.shape[0], dtype=bool)
    for drop in range(k + 1):
        kp = keys(numpy.delete(C, drop, axis=1))
        idx = numpy.minimum(numpy.searchsorted(feasible_keys, kp), len(feasible_keys) - 1)
        ok &= feasible_keys[idx] == kp
    return C[ok]


def inherit(C, parent_keys, parent_wit, slots):
    """(covered flags, inherited witness) of candidates C from the witnesses of their parents (slots: which to use)"""
    cm = to_masks(C)
    cov = numpy.zeros(C.shape[0], dtype=bool)
    got = numpy.zeros((C.shape[0], 2), dtype=numpy.uint64)
    for drop in range(C.shape[1] - 1, -1, -1):          # rows dropped from the highest down, like the kernel
        kp = keys(numpy.delete(C, drop, axis=1))
        idx = numpy.minimum(numpy.searchsorted(parent_keys, kp), len(parent_keys) - 1)
        found = parent_keys[idx] == kp
        for sl in slots:
            w = parent_wit[idx, sl]
            hit = found & ~cov & ((cm & ~w) == 0).all(1)
            got[hit] = w[hit]
            cov |= hit
    return cov, got


def main():
    sample = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    g = numpy.load(PATH)
    tw = Twin.from_npz(PATH)
    c3, st3 = g['level2_candidates'].astype(numpy.int16), g['level2_status']
    S3 = c3[(st3 & 3) == 3]
    t = time.time()
    cert3, piv3, wit3 = tw.k2w_witness(to_masks(S3), slots=2)
    print(f'level 3: {len(S3)} feasible candidates walked, {piv3} pivots, {cert3.mean():.4f} certified ({time.time() - t:.0f} s)', flush=True)
    S4 = children(S3, keys(S3))
    cov4, inh4 = inherit(S4, keys(S3), wit3, (0, 1))
    print(f'level 4: {len(S4)} candidates, {cov4.mean():.4f} inherit a certificate', flush=True)
    t = time.time()
    cert4, piv4, wit4 = tw.k2w_witness(to_masks(S4), slots=2, closed=cov4.astype(numpy.uint8))
    wit4[cov4, 0] = inh4[cov4]
    print(f'level 4: the walk certifies {cert4.sum()} of the {int((~cov4).sum())} others with {piv4} pivots ({time.time() - t:.0f} s)', flush=True)
    feas4 = cov4 | cert4.astype(bool)
    rng = numpy.random.default_rng(5)
    C = numpy.sort(numpy.array([rng.choice(M, size=5, replace=False) for _ in range(sample)]), axis=1).astype(numpy.int16)
    k4 = keys(S4[feas4])
    ok = numpy.ones(C.shape[0], dtype=bool)
    for drop in range(5):
        kp = keys(numpy.delete(C, drop, axis=1))
        idx = numpy.minimum(numpy.searchsorted(k4, kp), len(k4) - 1)
        ok &= k4[idx] == kp
    C = C[ok]
    for slots in ((0,), (0, 1)):
        cov5, _ = inherit(C, k4, wit4[feas4], slots)
        print(f'level 5 ({len(C)} sampled candidates): {cov5.mean():.4f} inherit with witness slot(s) {slots}', flush=True)


if __name__ == '__main__':
    main()
