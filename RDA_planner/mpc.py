"""`RDA_planner.mpc` — same import path as the reference (RDA_planner/mpc.py:12-15)."""
from rda_planner_b200.mpc import MPC, rdaobs  # noqa: F401
