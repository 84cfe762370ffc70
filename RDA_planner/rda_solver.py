"""`RDA_planner.rda_solver` — same import path as the reference (RDA_planner/rda_solver.py:17)."""
from rda_planner_b200.rda_solver import RDA_solver  # noqa: F401
