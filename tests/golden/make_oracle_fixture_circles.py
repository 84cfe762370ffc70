"""Generate tests/golden/oracle_circles_T30N20.npz: oracle trajectory for BASELINE configs[2] geometry
(T=30, 20 MOVING disc obstacles, min_sd=0.5, wu=0.2 as example/dynamic_obs/dynamic_obs.py:22) after
4 ADMM iterations.  Run in the build container (minutes)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from rda_planner_b200.scenarios import rectangle_robot, make_instance  # noqa: E402
from oracle.rda_oracle import OracleRDA  # noqa: E402

SEED, T, N, ITERS = 3000, 30, 20, 4
car = rectangle_robot(max_acce=(10, 1.0))
inst = make_instance(SEED, T=T, N=N, E=4, kind='circle', moving=True, lateral=(1.0, 6.0))
ref = [inst['ref'][:, t:t + 1] for t in range(T + 1)]
o = OracleRDA(T, car, max_edge_num=4, max_obs_num=N, iter_num=ITERS, iter_threshold=0.0, min_sd=0.5, wu=0.2)
u, info = o.iterative_solve(inst['nom_s'], inst['nom_u'], ref, inst['ref_speed'], list(inst['obstacles']))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_circles_T30N20.npz')
np.savez(out, seed=SEED, iters=ITERS, u=u, s=np.hstack(info['opt_state_list']), resi_dual=info['resi_dual'],
         resi_pri=info['resi_pri'])
print('wrote', out, o.cell_stats)
