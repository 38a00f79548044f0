"""Result container with the fields of the reference's CriticalRegion
(/root/reference/src/ppopt/critical_region.py:9-48).  Used only when PPOPT itself is not importable; when the solved
program is a genuine ppopt object the engine returns ppopt's own class (see mp_solvers/mpqp_combinatorial.py)."""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy


@dataclass(eq=False)
class CriticalRegion:
    r"""x(theta) = A theta + b,  lambda(theta) = C theta + d,  region = {theta : E theta <= f}."""
    A: numpy.ndarray
    b: numpy.ndarray
    C: numpy.ndarray
    d: numpy.ndarray
    E: numpy.ndarray
    f: numpy.ndarray
    active_set: List[int]

    omega_set: List[int] = field(default_factory=list)
    lambda_set: List[int] = field(default_factory=list)
    regular_set: List[List[int]] = field(default_factory=list)

    y_fixation: Optional[numpy.ndarray] = None
    y_indices: Optional[numpy.ndarray] = None
    x_indices: Optional[numpy.ndarray] = None

    def __repr__(self):
        return (f"Critical region with active set {self.active_set}\nThe Omega Constraint indices are {self.omega_set}"
                f"\nThe Lagrange multipliers Constraint indices are {self.lambda_set}"
                f"\nThe Regular Constraint indices are {self.regular_set}"
                f"\n A = {self.A} \n b = {self.b} \n C = {self.C} \n d = {self.d} \n E = {self.E} \n f = {self.f}")

    def evaluate(self, theta: numpy.ndarray) -> numpy.ndarray:
        if self.y_fixation is None:
            return self.A @ theta + self.b
        cont = self.A @ theta + self.b
        x_star = numpy.zeros((len(self.x_indices) + len(self.y_indices),))
        x_star[self.x_indices] = cont.flatten()
        x_star[self.y_indices] = self.y_fixation
        return x_star.reshape(-1, 1)

    def lagrange_multipliers(self, theta: numpy.ndarray) -> numpy.ndarray:
        return self.C @ theta + self.d

    def is_inside(self, theta: numpy.ndarray, tol: float = 1e-5) -> bool:
        return bool(numpy.all(self.E @ theta - self.f < tol))

    def get_constraints(self):
        return [self.E, self.f]
