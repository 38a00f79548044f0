"""Stand-in for `pathos` (not installed): ProcessingPool over multiprocess.Pool. TEST INFRASTRUCTURE."""
from . import multiprocessing  # noqa: F401
