"""Generate tests/golden/oracle_disc_robot.npz: float64 oracle traces (OracleRDA, every ADMM iteration) of planning instances
with a DISC body (car_tuple.cone_type 'norm2', /root/reference/RDA_planner/rda_solver.py:1034-1039).  The cells whose hinge is
active go through oracle/cell_generic.py (SLSQP in the original (lam, mu, z) variables with both second-order cones), which is
why the traces are committed instead of recomputed by the test suite.  Run in the build container:
    python tests/golden/make_oracle_fixture_disc_robot.py            (~15 min on 5 cores)
"""
import os
import sys
from multiprocessing import Pool

os.environ.setdefault('OMP_NUM_THREADS', '1')
os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
import numpy as np  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')
sys.path.insert(0, ROOT)
# name: seed, obstacle kind, moving, dynamics, body centre, radius, T, N, iterations
CASES = {
    'a': (61, 'polygon', False, 'diff', (0.0, 0.0), 1.2, 8, 4, 4),
    'b': (62, 'circle', False, 'omni', (0.0, 0.0), 1.2, 8, 4, 4),
    'c': (63, 'polygon', False, 'acker', (0.4, 0.0), 1.2, 8, 4, 4),
    'd': (64, 'polygon', False, 'diff', (0.0, 0.0), 1.0, 20, 8, 6),
    'e': (65, 'circle', True, 'acker', (0.3, -0.1), 0.9, 12, 6, 5),
}


def instance(name):
    from rda_planner_b200.scenarios import disc_robot, make_instance
    seed, kind, moving, dyn, center, radius, T, N, iters = CASES[name]
    car = disc_robot(radius=radius, center=center, wheelbase=2.0, dynamics=dyn)
    inst = make_instance(seed, T=T, N=N, E=4, lateral=(0.3, 3.0), kind=kind, moving=moving, dynamics=dyn)
    return car, inst, T, N, iters


def run(name):
    from oracle.rda_oracle import OracleRDA
    car, inst, T, N, iters = instance(name)
    ref = [inst['ref'][:, t:t + 1] for t in range(T + 1)]
    o = OracleRDA(T, car, max_edge_num=4, max_obs_num=N, iter_num=iters, iter_threshold=0.0)
    o.iterative_solve(inst['nom_s'], inst['nom_u'], ref, inst['ref_speed'], list(inst['obstacles']))
    tr = o.trace
    return name, dict(s=np.stack([x[0] for x in tr]), u=np.stack([x[1] for x in tr]), d=np.stack([x[4].reshape(-1) for x in tr]),
                      resi_dual=np.array([x[2] for x in tr]), resi_pri=np.array([x[3] for x in tr]), stats=str(o.cell_stats))


if __name__ == '__main__':
    with Pool(5) as pool:
        res = dict(pool.map(run, list(CASES), chunksize=1))
    flat = {}
    for name in CASES:
        for k in ('s', 'u', 'd', 'resi_dual', 'resi_pri'):
            flat[f'{name}_{k}'] = res[name][k]
        print(name, res[name]['stats'])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_disc_robot.npz')
    np.savez_compressed(out, **flat)
    print('wrote', out)
