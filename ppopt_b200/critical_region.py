"""Region record returned by the engine when PPOPT itself is not importable.

Field names and meaning follow the reference's result type (/root/reference/src/ppopt/critical_region.py:9-48) so that
code written against PPOPT reads the same attributes; when the solved program is a genuine ppopt object the engine
returns ppopt's own class instead (engine._region_classes)."""
import numpy


class CriticalRegion:
    """x*(theta) = A theta + b, lambda*(theta) = C theta + d on the polytope {theta : E theta <= f}."""

    __slots__ = ('A', 'b', 'C', 'd', 'E', 'f', 'active_set', 'omega_set', 'lambda_set', 'regular_set', 'y_fixation',
                 'y_indices', 'x_indices')

    def __init__(self, A, b, C, d, E, f, active_set, omega_set=None, lambda_set=None, regular_set=None, y_fixation=None,
                 y_indices=None, x_indices=None):
        self.A, self.b, self.C, self.d, self.E, self.f = A, b, C, d, E, f
        self.active_set = active_set
        self.omega_set = [] if omega_set is None else omega_set        # kept Theta rows
        self.lambda_set = [] if lambda_set is None else lambda_set     # constraints whose lambda >= 0 facet was kept
        self.regular_set = [] if regular_set is None else regular_set  # [positions, constraint indices] of kept inactive rows
        self.y_fixation, self.y_indices, self.x_indices = y_fixation, y_indices, x_indices

    def __repr__(self):
        return (f'CriticalRegion(active_set={self.active_set}, omega={self.omega_set}, lambda={self.lambda_set}, '
                f'regular={self.regular_set}, E{numpy.shape(self.E)}, A{numpy.shape(self.A)})')

    def evaluate(self, theta):
        x = self.A @ theta + self.b
        if self.y_fixation is None:
            return x
        full = numpy.zeros(len(self.x_indices) + len(self.y_indices))
        full[self.x_indices] = x.ravel()
        full[self.y_indices] = self.y_fixation
        return full.reshape(-1, 1)

    def lagrange_multipliers(self, theta):
        return self.C @ theta + self.d

    def is_inside(self, theta, tol=1e-5) -> bool:
        return bool((self.E @ theta - self.f < tol).all())

    def is_full_dimension(self) -> bool:
        """Chebyshev radius of {theta : E theta <= f} > 1e-8 (critical_region.py:89-105), one small LP on the GPU; use
        ppopt_b200.chebyshev.full_dimensional(regions) to test many regions in one launch"""
        from .chebyshev import RADIUS_TOL, chebyshev_radii
        return bool(chebyshev_radii([(self.E, self.f)])[0] > RADIUS_TOL)

    def get_constraints(self):
        return [self.E, self.f]
