"""Generate tests/golden/oracle_configs.npz: oracle traces of ONE instance of the BASELINE configs at their real sizes —
C (T=30, 20 moving discs, min_sd=0.5, wu=0.2), D (T=30, N=64 hulls with up to 8 vertices, slack_gain=13) and
E (T=40, N=128 polytopes with up to 8 faces) — for a few ADMM iterations (the float64 numpy oracle needs minutes per
iteration at the E size: dense interior point on 5 000 hinge slacks).  Run in the build container:
    python tests/golden/make_oracle_fixture_configs.py            (~40 min on 3 cores)
"""
import os
import sys
from multiprocessing import Pool

os.environ.setdefault('OMP_NUM_THREADS', '2')
os.environ.setdefault('OPENBLAS_NUM_THREADS', '2')
import numpy as np  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')
sys.path.insert(0, ROOT)
CASES = [('C', 7000, 4), ('D', 7000, 3), ('E', 7000, 2)]


def run(case):
    from rda_planner_b200.scenarios import rectangle_robot, config_instance, CONFIGS
    from oracle.rda_oracle import OracleRDA
    name, seed, iters = case
    c = CONFIGS[name]
    inst = config_instance(name, seed)
    T = c['T']
    ref = [inst['ref'][:, t:t + 1] for t in range(T + 1)]
    o = OracleRDA(T, rectangle_robot(dynamics=c['dynamics']), max_edge_num=c['E'], max_obs_num=c['N'], iter_num=iters,
                  iter_threshold=0.0, **c['tun'])
    o.iterative_solve(inst['nom_s'], inst['nom_u'], ref, inst['ref_speed'], list(inst['obstacles']))
    tr = o.trace
    return name, dict(s=np.stack([x[0] for x in tr]), u=np.stack([x[1] for x in tr]), d=np.stack([x[4].reshape(-1) for x in tr]),
                      resi_dual=np.array([x[2] for x in tr]), resi_pri=np.array([x[3] for x in tr]), stats=str(o.cell_stats))


if __name__ == '__main__':
    with Pool(3) as pool:
        res = dict(pool.map(run, CASES, chunksize=1))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_configs.npz')
    flat = {}
    for name, seed, iters in CASES:
        flat[f'{name}_seed'] = seed
        flat[f'{name}_iters'] = iters
        for k in ('s', 'u', 'd', 'resi_dual', 'resi_pri'):
            flat[f'{name}_{k}'] = res[name][k]
        print(name, res[name]['stats'])
    np.savez_compressed(out, **flat)
    print('wrote', out)
