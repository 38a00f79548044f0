"""Host-side pruning bookkeeping with the reference's API (/root/reference/src/ppopt/mp_solvers/solver_utils.py:15-55,154-166).

The GPU path does its pruning on bitmasks (csrc/k6_children.cu); these two symbols are kept because user code and the
reference's tests call them directly.  They are plain set arithmetic - no numerics."""
from typing import List


class CombinationTester:
    """Remembers infeasible active-set combinations and rejects their supersets."""

    def __init__(self):
        self.combos = set()
        self.new_combos = set()

    def check(self, active_set) -> bool:
        """False if ``active_set`` contains a stored infeasible combination (it can be culled), True otherwise."""
        cand = active_set if isinstance(active_set, set) else set(active_set)
        if not cand:
            return True
        return not any(cand.issuperset(c) for c in self.combos)

    def add_combo(self, active_set) -> None:
        if isinstance(active_set, set):
            return  # the reference ignores plain sets (solver_utils.py:48-52)
        self.combos.add(tuple(active_set))

    def add_combos(self, set_list) -> None:
        self.combos.update(set_list)


def generate_children_sets(active_set, num_constraints: int, murder_list=None) -> List[List[int]]:
    """All supersets of cardinality + 1 that extend ``active_set`` past its last index and survive the pruning list."""
    ok = (lambda x: True) if murder_list is None else murder_list.check
    start = 0 if len(active_set) == 0 else active_set[-1] + 1
    return [[*active_set, i] for i in range(start, num_constraints) if ok([*active_set, i])]


def _accept(attempted, murder_list):
    """neighbour filter of the graph algorithms (solver_utils.py:56-66): not attempted yet, not a superset of a pruned set"""
    if attempted is None:
        return (lambda _: True) if murder_list is None else murder_list.check
    if murder_list is None:
        return lambda x: x not in attempted
    return lambda x: x not in attempted and murder_list.check(x)


def generate_reduce(candidate: tuple, murder_list=None, attempted=None, equality_set=None) -> list:
    """every active set with one constraint of ``candidate`` dropped that keeps all equalities (solver_utils.py:69-83)"""
    equality_set = set() if equality_set is None else equality_set
    ok = _accept(attempted, murder_list)
    out = []
    for i in candidate:
        possible = tuple(sorted(j for j in candidate if j != i))
        if ok(possible) and set(possible).issuperset(equality_set):
            out.append(possible)
    return out


def generate_extra(candidate: tuple, expansion_set, murder_list=None, attempted=None) -> list:
    """``candidate`` plus one constraint of ``expansion_set`` (the facets of its region), filtered (solver_utils.py:86-107)"""
    ok = _accept(attempted, murder_list)
    out = []
    for c in expansion_set:
        child = tuple(sorted([*candidate, c]))
        if ok(child):
            out.append(child)
    return out
