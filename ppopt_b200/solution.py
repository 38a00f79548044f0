"""Solution container with the surface of the reference's Solution (/root/reference/src/ppopt/solution.py:15-112)."""
from typing import List, Optional

import numpy

from .critical_region import CriticalRegion


class Solution:
    def __init__(self, program, critical_regions: List[CriticalRegion], is_overlapping=False,
                 point_location_tolerance=1e-5):
        self.program = program
        self.critical_regions = critical_regions
        self.is_overlapping = is_overlapping
        self.point_location_tolerance = point_location_tolerance

    def add_region(self, region: CriticalRegion) -> None:
        self.critical_regions.append(region)

    def evaluate(self, theta_point: numpy.ndarray) -> Optional[numpy.ndarray]:
        cr = self.get_region(theta_point)
        return None if cr is None else cr.evaluate(theta_point)

    def get_region(self, theta_point: numpy.ndarray) -> Optional[CriticalRegion]:
        if self.is_overlapping:
            return self.get_region_overlap(theta_point)
        return self.get_region_no_overlap(theta_point)

    def get_region_no_overlap(self, theta_point):
        for region in self.critical_regions:
            if region.is_inside(theta_point, self.point_location_tolerance):
                return region
        return None

    def get_region_overlap(self, theta_point):
        best, best_obj = None, float('inf')
        for region in self.critical_regions:
            if region.is_inside(theta_point, self.point_location_tolerance):
                obj = self.program.evaluate_objective(region.evaluate(theta_point), theta_point)
                if obj <= best_obj:
                    best, best_obj = region, obj
        return best

    def __len__(self):
        return len(self.critical_regions)
