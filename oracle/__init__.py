"""TEST INFRASTRUCTURE ONLY — CPU (float64) restatement of the RDA-planner ADMM hot path.

PARITY UNPINNED: the reference's solver stack (cvxpy 1.5.2 / ECOS / pathos) is not
installable in this image and the reference ships no golden vectors; see
oracle/rda_oracle.py for what pins this oracle instead.

Nothing in rda_planner_b200/ or RDA_planner/ (the product) may import this package;
only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs do.
"""
