"""Generate tests/golden/oracle_metric50.npz: per-iteration oracle traces of the BASELINE metric row
(T=30, N=20 polygons, E=4, 50 ADMM iterations, iter_threshold=0) for

  * 16 instances of the bench workload (seeds 9000..9015, the first instances bench.py builds), and
  * 8 instances of the harsher generator the parity tests use (lateral=(0.3, 3.5), seeds 1000..1007:
    obstacles on the path, overlap cells).

Stored per instance and iteration: s (3,T+1), u (2,T), d (T), resi_dual, resi_pri and the literal su cost
(total / nav / hinge / consensus, OracleRDA.su_cost_literal).  Run in the build container:
    python tests/golden/make_oracle_fixture_50.py [procs]       (~25 min on 6 cores)
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')
sys.path.insert(0, ROOT)
T, N, E, ITERS = 30, 20, 4, 50
CASES = [(9000 + i, (1.8, 6.0)) for i in range(16)] + [(1000 + i, (0.3, 3.5)) for i in range(8)]


def run(case):
    from rda_planner_b200.scenarios import rectangle_robot, make_instance
    from oracle.rda_oracle import OracleRDA
    seed, lateral = case
    inst = make_instance(seed, T=T, N=N, E=E, lateral=lateral)
    ref = [inst['ref'][:, t:t + 1] for t in range(T + 1)]
    o = OracleRDA(T, rectangle_robot(), max_edge_num=E, max_obs_num=N, iter_num=ITERS, iter_threshold=0.0)
    o.iterative_solve(inst['nom_s'], inst['nom_u'], ref, inst['ref_speed'], list(inst['obstacles']))
    tr = o.trace
    return dict(
        s=np.stack([x[0] for x in tr]), u=np.stack([x[1] for x in tr]), d=np.stack([x[4].reshape(-1) for x in tr]),
        resi_dual=np.array([x[2] for x in tr]), resi_pri=np.array([x[3] for x in tr]),
        cost=np.array([[x[5][k] for k in ('total', 'nav', 'hinge', 'consensus')] for x in tr]),
        stats=str(o.cell_stats))


if __name__ == '__main__':
    procs = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    with Pool(procs) as pool:
        res = pool.map(run, CASES, chunksize=1)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_metric50.npz')
    np.savez_compressed(
        out, seeds=np.array([c[0] for c in CASES]), lateral=np.array([c[1] for c in CASES]), iters=ITERS,
        s=np.stack([r['s'] for r in res]), u=np.stack([r['u'] for r in res]), d=np.stack([r['d'] for r in res]),
        resi_dual=np.stack([r['resi_dual'] for r in res]), resi_pri=np.stack([r['resi_pri'] for r in res]),
        cost=np.stack([r['cost'] for r in res]), cost_keys=np.array(['total', 'nav', 'hinge', 'consensus']))
    print('wrote', out)
    for c, r in zip(CASES, res):
        print(c, r['stats'])
