"""ctypes binding of oracle/_build/libtwin.so (the CPU checker of the batched algorithms).  TEST INFRASTRUCTURE."""
import ctypes
import os
import subprocess

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(HERE, '_build', 'libtwin.so')
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.check_call(['make', '-C', HERE])
        l = ctypes.CDLL(_PATH)
        l.twin_create.restype = ctypes.c_void_p
        l.twin_create.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 8
        l.twin_destroy.argtypes = [ctypes.c_void_p]
        l.twin_words.argtypes = [ctypes.c_void_p]
        l.twin_use_gram.argtypes = [ctypes.c_void_p]
        l.twin_pivots.argtypes = [ctypes.c_void_p]
        l.twin_pivots.restype = ctypes.c_long
        l.twin_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        l.twin_emit.argtypes = [ctypes.c_void_p] * 6
        l.twin_emit_margins.argtypes = [ctypes.c_void_p] * 7
        l.twin_k2a.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 3
        l.twin_k2p.argtypes = l.twin_k2a.argtypes
        l.twin_feas_rhs.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib = l
    return _lib


def _c(a):
    return numpy.ascontiguousarray(numpy.asarray(a, dtype=numpy.float64))


class Twin:
    def __init__(self, A, b, F, A_t, b_t, Q, c, H, n_eq, is_qp):
        A, F = _c(A), _c(F)
        self.n, self.t, self.m = A.shape[1], F.shape[1], A.shape[0]
        A_t = _c(A_t).reshape(-1, self.t)
        self.q, self.n_eq, self.is_qp = A_t.shape[0], int(n_eq), bool(is_qp)
        Q = _c(Q) if is_qp else numpy.zeros((self.n, self.n))
        self._keep = [A, _c(b).ravel(), F, A_t, _c(b_t).ravel(), Q, _c(c).ravel(), _c(H)]
        self.h = lib().twin_create(self.n, self.t, self.m, self.q, self.n_eq, int(self.is_qp), *[x.ctypes.data for x in self._keep])
        if not self.h:
            raise RuntimeError('twin_create failed')
        self.W = lib().twin_words(self.h)
        self.R0 = self.m - self.n_eq + self.q

    @classmethod
    def from_npz(cls, path):
        g = numpy.load(path)
        is_qp = str(g['kind']) == 'qp'
        return cls(g['A'], g['b'], g['F'], g['A_t'], g['b_t'], g['Q'] if is_qp else None, g['c'], g['H'], int(g['n_eq']), is_qp)

    def masks(self, active_sets):
        out = numpy.zeros((len(active_sets), self.W), dtype=numpy.uint64)
        for ci, aset in enumerate(active_sets):
            for idx in aset:
                i = int(idx) - self.n_eq
                if i >= 0:
                    out[ci, i >> 6] |= numpy.uint64(1) << numpy.uint64(i & 63)
        return out

    def eval(self, masks, aux=False):
        masks = numpy.ascontiguousarray(masks).view(numpy.uint64).reshape(-1, self.W)
        st = numpy.zeros(masks.shape[0], dtype=numpy.uint8)
        ax = numpy.zeros((masks.shape[0], 3))
        lib().twin_eval(self.h, masks.ctypes.data, masks.shape[0], 0, st.ctypes.data, ax.ctypes.data)
        return (st, ax) if aux else st

    def emit(self, mask, margins=False):
        mask = numpy.ascontiguousarray(mask).view(numpy.uint64).reshape(self.W)
        k = self.n_eq + int(sum(bin(int(w)).count('1') for w in mask))
        laws = numpy.zeros((self.n + k, self.t + 1))
        rows = numpy.zeros((self.R0, self.t + 1))
        flags = numpy.zeros(self.R0, dtype=numpy.int32)
        info = numpy.zeros(4)
        if margins:
            mg = numpy.full(self.R0, numpy.nan)
            rc = lib().twin_emit_margins(self.h, mask.ctypes.data, laws.ctypes.data, rows.ctypes.data, flags.ctypes.data,
                                         info.ctypes.data, mg.ctypes.data)
            return rc, laws, rows, flags, info, mg
        rc = lib().twin_emit(self.h, mask.ctypes.data, laws.ctypes.data, rows.ctypes.data, flags.ctypes.data, info.ctypes.data)
        return rc, laws, rows, flags, info

    def k2a(self, masks, max_iter=96, max_iter2=96, prefix=False):
        """sequential K2a (relaxation certificates): (certified flags, steps, exact residuals of the uncertified ones);
        prefix=True: the prefix-projected form of csrc/k2p_prefix.cu instead of the per-candidate form of k2a_relax.cu"""
        masks = numpy.ascontiguousarray(masks).view(numpy.uint64).reshape(-1, self.W)
        n = masks.shape[0]
        flags, steps = numpy.zeros(n, dtype=numpy.int32), numpy.zeros(n, dtype=numpy.int32)
        resid = numpy.zeros((n, self.R0))
        (lib().twin_k2p if prefix else lib().twin_k2a)(self.h, masks.ctypes.data, n, int(max_iter), int(max_iter2),
                                                       flags.ctypes.data, steps.ctypes.data, resid.ctypes.data)
        return flags, steps, resid

    def feas_from(self, ineq_rows, rhs=None):
        """feasibility LP of one active set (inequality-row indices), optionally with a replacement rhs column (the LP seen
        from another origin, as K2 starts from K2a's last iterate): (feasible, pivots)"""
        act = numpy.ascontiguousarray(ineq_rows, dtype=numpy.int32)
        piv = ctypes.c_int(0)
        r = None if rhs is None else numpy.ascontiguousarray(rhs, dtype=numpy.float64)
        f = lib().twin_feas_rhs(self.h, act.ctypes.data, len(act), None if r is None else r.ctypes.data, ctypes.byref(piv))
        return bool(f), piv.value

    def walk_dict(self):
        """start vertex of the K2w walk: (D0 [beta | D], basic rows, nonbasic rows, T0) or None when the program has none"""
        l = lib()
        l.twin_walk_dict.argtypes = [ctypes.c_void_p] * 6
        l.twin_nfree.argtypes = [ctypes.c_void_p]
        l.twin_rows.argtypes = [ctypes.c_void_p]
        l.twin_t0.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        nb, ld = ctypes.c_int(0), ctypes.c_int(0)
        if not l.twin_walk_dict(self.h, ctypes.byref(nb), ctypes.byref(ld), None, None, None):
            return None
        nf, R0 = l.twin_nfree(self.h), l.twin_rows(self.h)
        D0 = numpy.zeros((nb.value, ld.value))
        bvar, nvar = numpy.zeros(nb.value, dtype=numpy.int32), numpy.zeros(nf, dtype=numpy.int32)
        l.twin_walk_dict(self.h, ctypes.byref(nb), ctypes.byref(ld), D0.ctypes.data, bvar.ctypes.data, nvar.ctypes.data)
        T0 = numpy.zeros((R0, nf + 2))
        l.twin_t0(self.h, T0.ctypes.data)
        return D0, bvar, nvar, T0

    def k2w(self, masks):
        """sequential K2w (vertex walk) over one level in lexicographic order: (certified flags, pivots)"""
        l = lib()
        l.twin_k2w.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
        l.twin_k2w.restype = ctypes.c_long
        masks = numpy.ascontiguousarray(masks).view(numpy.uint64).reshape(-1, self.W)
        cert = numpy.zeros(masks.shape[0], dtype=numpy.uint8)
        piv = l.twin_k2w(self.h, masks.ctypes.data, masks.shape[0], cert.ctypes.data)
        return cert, int(piv)

    def k2w_witness(self, masks, slots=1, closed=None):
        """sequential K2w with witnesses: (certified flags, pivots, witness masks).  The witness of a certified candidate is
        the set of all rows active at the certifying vertex; with slots >= 2, slot 1 holds a later vertex of the walk that
        holds the candidate as well (the kernel's revisits).  slots == 1: witness is n x ceil(R0 / 64) uint64, else
        n x slots x ceil(R0 / 64).  closed: optional flags of candidates that arrive certified (inherited) - part of their
        segment, never a target of the walk, revisited like the others."""
        l = lib()
        l.twin_k2w_witness.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_void_p]
        l.twin_k2w_witness.restype = ctypes.c_long
        masks = numpy.ascontiguousarray(masks).view(numpy.uint64).reshape(-1, self.W)
        cert = numpy.zeros(masks.shape[0], dtype=numpy.uint8)
        wit = numpy.zeros((masks.shape[0], slots, (self.R0 + 63) // 64), dtype=numpy.uint64)
        cl = None if closed is None else numpy.ascontiguousarray(closed, dtype=numpy.uint8)
        piv = l.twin_k2w_witness(self.h, masks.ctypes.data, masks.shape[0], cert.ctypes.data, wit.ctypes.data, slots,
                                 None if cl is None else cl.ctypes.data)
        return cert, int(piv), (wit[:, 0] if slots == 1 else wit)

    def pivots(self):
        return lib().twin_pivots(self.h)

    def __del__(self):
        try:
            if self.h:
                lib().twin_destroy(self.h)
        except Exception:
            pass
