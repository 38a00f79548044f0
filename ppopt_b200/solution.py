"""Solution record with the surface of the reference's Solution (/root/reference/src/ppopt/solution.py:15-112): the
solved program, the ordered list of critical regions, and first-hit / best-objective point location."""


class Solution:
    def __init__(self, program, critical_regions, is_overlapping=False, point_location_tolerance=1e-5):
        self.program = program
        self.critical_regions = critical_regions
        self.is_overlapping = is_overlapping
        self.point_location_tolerance = point_location_tolerance

    def add_region(self, region):
        self.critical_regions.append(region)

    def __len__(self):
        return len(self.critical_regions)

    def _containing(self, theta):
        tol = self.point_location_tolerance
        return (r for r in self.critical_regions if r.is_inside(theta, tol))

    def get_region_no_overlap(self, theta):
        return next(self._containing(theta), None)

    def get_region_overlap(self, theta):
        best, best_val = None, float('inf')
        for r in self._containing(theta):
            val = self.program.evaluate_objective(r.evaluate(theta), theta)
            if val <= best_val:
                best, best_val = r, val
        return best

    def get_region(self, theta):
        return self.get_region_overlap(theta) if self.is_overlapping else self.get_region_no_overlap(theta)

    def evaluate(self, theta):
        r = self.get_region(theta)
        return None if r is None else r.evaluate(theta)
