"""Measure what the tie-break of the non-unique LamMuZ argmin changes.

Runs the float64 oracle twice on the same metric-row instances: with the shipped rule (max margin,
LP-vertex multipliers, z = stuff/2; oracle/cell_geo.py) and with the interior-point rule an ECOS-like
solver follows (analytic centre of the optimal face, oracle/cell_ac.py), and reports the gap of
trajectories, residuals and early-stop iteration.  CPU only; ~1 min per instance and 10 iterations.

    python tools/tiebreak_gap.py [instances=16] [iterations=20] [procs=6]
"""
import json
import os
import sys
from multiprocessing import Pool

os.environ.setdefault('OMP_NUM_THREADS', '1')
os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
import numpy as np  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
T, N, E = 30, 20, 4


def run(args):
    seed, iters = args
    from rda_planner_b200.scenarios import rectangle_robot, make_instance
    from oracle.rda_oracle import OracleRDA
    out = {}
    for rule in ('geo', 'ac'):
        inst = make_instance(seed, T=T, N=N, E=E)
        ref = [inst['ref'][:, t:t + 1] for t in range(T + 1)]
        o = OracleRDA(T, rectangle_robot(), max_edge_num=E, max_obs_num=N, iter_num=iters, iter_threshold=0.0,
                      cell_solver=rule)
        o.iterative_solve(inst['nom_s'], inst['nom_u'], ref, inst['ref_speed'], list(inst['obstacles']))
        out[rule] = dict(s=np.stack([x[0] for x in o.trace]), u=np.stack([x[1] for x in o.trace]),
                         rd=np.array([x[2] for x in o.trace]), rp=np.array([x[3] for x in o.trace]),
                         nav=np.array([x[5]['nav'] for x in o.trace]))
    return out


if __name__ == '__main__':
    ninst = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    procs = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    with Pool(procs) as pool:
        res = pool.map(run, [(9000 + i, iters) for i in range(ninst)], chunksize=1)
    rep = {'instances': ninst, 'iterations': iters, 'per_iteration': []}
    thr = 0.2

    def first_stop(r):
        ok = np.nonzero((r['rd'] < thr) & (r['rp'] < thr))[0]
        return int(ok[0]) + 1 if ok.size else iters + 1
    for k in [0, 1, 3, 7, iters - 1]:
        if k >= iters:
            continue
        ds = np.array([np.abs(r['geo']['s'][k][:2] - r['ac']['s'][k][:2]).max() for r in res])
        du = np.array([np.abs(r['geo']['u'][k] - r['ac']['u'][k]).max() for r in res])
        rep['per_iteration'].append({
            'iteration': k + 1, 'pos_gap_m': {'median': float(np.median(ds)), 'p90': float(np.quantile(ds, .9)), 'max': float(ds.max())},
            'u_gap': {'median': float(np.median(du)), 'max': float(du.max())},
            'resi_dual': {'geo_median': float(np.median([r['geo']['rd'][k] for r in res])),
                          'ac_median': float(np.median([r['ac']['rd'][k] for r in res]))},
            'resi_pri': {'geo_median': float(np.median([r['geo']['rp'][k] for r in res])),
                         'ac_median': float(np.median([r['ac']['rp'][k] for r in res]))},
            'nav_cost_rel_gap_median': float(np.median([abs(r['geo']['nav'][k] - r['ac']['nav'][k]) / max(abs(r['geo']['nav'][k]), 1e-9) for r in res]))})
    rep['early_stop_iteration(thr=0.2)'] = {'geo': [first_stop(r['geo']) for r in res], 'ac': [first_stop(r['ac']) for r in res]}
    print(json.dumps(rep, indent=1))
    with open(os.path.join(ROOT, 'profiles', 'tiebreak_gap_r02.json'), 'w') as f:
        json.dump(rep, f, indent=1)
