"""Drop-in for ppopt.mp_solvers.mpqp_combinatorial (/root/reference/src/ppopt/mp_solvers/mpqp_combinatorial.py).

``solve(program) -> Solution`` has the reference's signature and return type; the whole body runs on the GPU
(ppopt_b200.engine).  ``check_child_feasibility`` keeps its reference signature for callers/tests that use it directly
(tests/mpqp_solver_tests/test_mpqp_combinatorial.py:85-87).
"""
from typing import List

from .. import engine as _engine
from .._lib import ST_FEAS
from .solver_utils import CombinationTester


def solve(program):
    """Solves the mpQP / mpLP with the combinatorial (Gupta et al. 2011) enumeration, on the GPU."""
    return _engine.solve(program)


def check_child_feasibility(program, set_list: List[List[int]], combination_checker: CombinationTester) -> List[List[int]]:
    """Feasible members of ``set_list``; the infeasible ones are added to ``combination_checker``
    (mpqp_combinatorial.py:75-92).  One batched K1+K2 launch per distinct cardinality."""
    eng = _engine.Engine(_engine.program_arrays(program))
    try:
        feasible = [False] * len(set_list)
        by_k = {}
        for pos, s in enumerate(set_list):
            # kernels are specialised on the number of activated INEQUALITY rows (the bits of the mask); a set that omits
            # some equality rows is evaluated as its closure (equalities are always active, mplp_program.py:112-118)
            by_k.setdefault(sum(1 for i in s if int(i) >= eng.n_eq), []).append(pos)
        for k_act, positions in by_k.items():
            masks = eng.masks_from_lists([set_list[p] for p in positions])
            st = eng.level_eval(masks, k_act, stages=3).cpu().numpy()
            for p, s in zip(positions, st):
                feasible[p] = bool(s & ST_FEAS)
    finally:
        eng.close()
    out = []
    for s, ok in zip(set_list, feasible):
        if ok:
            out.append(s)
        else:
            combination_checker.add_combo(s)
    return out
