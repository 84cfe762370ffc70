"""Drop-in package name of the reference (hanruihua/RDA-planner): `from RDA_planner.mpc import MPC`
keeps working; the implementation lives in rda_planner_b200 (B200 CUDA kernels)."""
