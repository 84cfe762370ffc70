"""rda_planner_b200 — B200-native ADMM-MPC hot path behind the RDA-planner API."""
