"""Generate tests/golden/oracle_halfspace.npz: float64 oracle traces of instances whose obstacles are GENERAL half-space sets
handed over as (A, b) tuples (MPC(rda_obstacle=True), /root/reference/RDA_planner/mpc.py:150-155): two corridor walls (one
row each), a wedge (two rows) and a box whose rows are shuffled and carry a redundant fifth row.  The oracle solves every
cell in the ORIGINAL rows with oracle/cell_generic.py (SLSQP + HiGHS, no vertex geometry), i.e. it never sees the closing
square rda_planner_b200.rda_solver.canonical_polygon_rows adds.  Run in the build container:
    python tests/golden/make_oracle_fixture_halfspace.py            (~10 min)
"""
import os
import sys
from multiprocessing import Pool

os.environ.setdefault('OMP_NUM_THREADS', '1')
os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
import numpy as np  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')
sys.path.insert(0, ROOT)
CASES = {'w': (71, 'diff', 10, 4, 4), 'x': (72, 'acker', 8, 4, 4)}      # name: seed, dynamics, T, N (= obstacles), iterations


def instance(name):
    """Robot driving along a straight line through a corridor: walls at +-half-width, a wedge poking in from the left
    ahead of the robot, a box with shuffled + redundant rows on the right."""
    from rda_planner_b200.scenarios import rectangle_robot, rollout
    from rda_planner_b200.mpc import rdaobs, polygon_halfspaces
    seed, dyn, T, N, iters = CASES[name]
    rng = np.random.default_rng(seed)
    heading = rng.uniform(-np.pi, np.pi)
    d = np.array([np.cos(heading), np.sin(heading)])
    n = np.array([-d[1], d[0]])
    start = rng.uniform(10, 40, 2)
    state = np.array([start[0] + 0.4 * n[0], start[1] + 0.4 * n[1], heading + 0.05])
    v = 4.0
    nom_u = np.vstack([np.full(T, v), np.zeros(T)])
    nom_s = rollout(state, nom_u, 0.1, 3.0, dyn)
    ref = np.zeros((3, T + 1))
    for t in range(T + 1):
        ref[0:2, t] = start + d * (v * 0.1 * t)
        ref[2, t] = heading
    half = 2.2
    obs = []
    # walls: n.x >= n.start + half  (left), n.x <= n.start - half (right) are the OBSTACLES
    obs.append(rdaobs(-n[None] * 1.7, np.array([[-(n @ start + half) * 1.7]]), 'Rpositive', None, None))
    obs.append(rdaobs(n[None] * 0.6, np.array([[(n @ start - half) * 0.6]]), 'Rpositive', None, None))
    # wedge with its apex 1.3 m left of the line, 3 m ahead, opening away from the corridor
    apex = start + d * 3.0 + n * 1.3
    a1, a2 = heading + np.pi / 2 + 0.9, heading + np.pi / 2 - 0.9
    Aw = -np.array([[np.cos(a1), np.sin(a1)], [np.cos(a2), np.sin(a2)]])
    obs.append(rdaobs(Aw, (Aw @ apex)[:, None], 'Rpositive', None, None))
    # box on the right, rows shuffled, plus a redundant row
    c = start + d * 2.0 - n * 1.9
    V = np.array([c + d * 1.0 + n * 0.5, c - d * 1.0 + n * 0.5, c - d * 1.0 - n * 0.5, c + d * 1.0 - n * 0.5]).T
    Ab, bb = polygon_halfspaces(V)
    perm = [2, 0, 3, 1]
    Ab = np.vstack([Ab[perm], [[d[0] + n[0], d[1] + n[1]]]])
    bb = np.concatenate([np.ravel(bb)[perm], [(d + n) @ c + 50.0]])[:, None]
    obs.append(rdaobs(Ab, bb, 'Rpositive', None, None))
    car = rectangle_robot(dynamics=dyn)
    return car, dict(nom_s=nom_s, nom_u=nom_u, ref=ref, ref_speed=v, obstacles=obs[:N]), T, N, iters


def run(name):
    from oracle.rda_oracle import OracleRDA
    car, inst, T, N, iters = instance(name)
    ref = [inst['ref'][:, t:t + 1] for t in range(T + 1)]
    o = OracleRDA(T, car, max_edge_num=5, max_obs_num=N, iter_num=iters, iter_threshold=0.0, cell_solver='generic')
    o.iterative_solve(inst['nom_s'], inst['nom_u'], ref, inst['ref_speed'], list(inst['obstacles']))
    tr = o.trace
    return name, dict(s=np.stack([x[0] for x in tr]), u=np.stack([x[1] for x in tr]), d=np.stack([x[4].reshape(-1) for x in tr]),
                      resi_dual=np.array([x[2] for x in tr]), resi_pri=np.array([x[3] for x in tr]))


if __name__ == '__main__':
    with Pool(2) as pool:
        res = dict(pool.map(run, list(CASES), chunksize=1))
    flat = {}
    for name in CASES:
        for k in ('s', 'u', 'd', 'resi_dual', 'resi_pri'):
            flat[f'{name}_{k}'] = res[name][k]
        print(name, res[name]['resi_dual'], res[name]['resi_pri'])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_halfspace.npz')
    np.savez_compressed(out, **flat)
    print('wrote', out)
