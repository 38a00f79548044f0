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
