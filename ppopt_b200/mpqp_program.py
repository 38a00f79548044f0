from .mplp_program import MPLP_Program, MPQP_Program  # noqa: F401  (same import path as ppopt.mpqp_program)
