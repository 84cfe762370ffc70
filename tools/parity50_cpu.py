"""Distribution of the gaps between the CPU build of the kernel cores (oracle/cpu_port) and an oracle trace at the benchmarked
configuration (default: tests/golden/oracle_metric50.npz), by ADMM iteration — the CPU twin of tests/test_gpu_parity50.py.
    python tools/parity50_cpu.py [fixture.npz] [out.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from oracle import cpu_port  # noqa: E402
from rda_planner_b200.rda_solver import pack_obstacles  # noqa: E402
from rda_planner_b200.scenarios import rectangle_robot, make_instance  # noqa: E402

T, N, E = 30, 20, 4
CHECK = [1, 2, 4, 8, 16, 32, 50]


def main(fixture, out=None):
    z = np.load(fixture)
    B = len(z['seeds'])
    car = rectangle_robot()
    insts = [make_instance(int(sd), T=T, N=N, E=E, lateral=tuple(l)) for sd, l in zip(z['seeds'], z['lateral'])]
    packs = [pack_obstacles(list(i['obstacles']), T, N, E) for i in insts]
    st = lambda k: np.stack([i[k] for i in insts]).astype(np.float32)
    q = lambda x: {'median': float(np.median(x)), 'p75': float(np.quantile(x, .75)), 'p95': float(np.quantile(x, .95)), 'max': float(x.max())}
    rows = []
    for it in CHECK:
        r = cpu_port.solve_batch(car, T, N, E, st('nom_s'), st('nom_u'), st('ref'), np.array([i['ref_speed'] for i in insts], np.float32),
                                 np.stack([p[0] for p in packs]), np.stack([p[1] for p in packs]), np.stack([p[2] for p in packs]),
                                 np.array([p[3] for p in packs], np.int32), iter_num=it, iter_threshold=0.0)
        ds = np.abs(r['s'] - z['s'][:, it - 1]).reshape(B, -1).max(1)
        du = np.abs(r['u'] - z['u'][:, it - 1]).reshape(B, -1).max(1)
        rp = np.abs(r['resi_pri'] - z['resi_pri'][:, it - 1]) / (1 + z['resi_pri'][:, it - 1])
        rd = np.abs(r['resi_dual'] - z['resi_dual'][:, it - 1]) / (1 + z['resi_dual'][:, it - 1])
        rows.append({'iteration': it, 'state_gap': q(ds), 'control_gap': q(du), 'resi_pri_rel_gap': q(rp), 'resi_dual_rel_gap': q(rd),
                     'state_gap_bench16': q(ds[:16]), 'state_gap_harsh8': q(ds[16:])})
        print(f"it {it:2d}: |ds| med {np.median(ds):.1e} p75 {np.quantile(ds, .75):.1e} p95 {np.quantile(ds, .95):.1e} max {ds.max():.1e}   "
              f"|du| med {np.median(du):.1e} max {du.max():.1e}   resi_pri med {np.median(rp):.1e} resi_dual med {np.median(rd):.1e}")
    if out:
        with open(out, 'w') as f:
            json.dump({'what': 'CPU build of the kernel cores (oracle/cpu_port: float32 state, float64 su-QP) against the oracle trace',
                       'fixture': os.path.relpath(fixture, ROOT), 'instances': B, 'rows': rows}, f, indent=1)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'tests', 'golden', 'oracle_metric50.npz'),
         sys.argv[2] if len(sys.argv) > 2 else None)
