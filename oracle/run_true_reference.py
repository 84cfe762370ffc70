"""TEST INFRASTRUCTURE — pin the oracle against the TRUE reference wherever cvxpy + ECOS exist.

This image has neither, so parity of the solver path is "unpinned": the oracle is a
restatement, checked piecewise (tests/test_oracle.py) and against the reference's executable numpy
helpers (oracle/gen_golden.py).  In an environment where `import cvxpy, ecos` works and the reference
checkout is available, this script closes the gap:

    python oracle/run_true_reference.py --ref /path/to/RDA-planner [--write]

It runs the unmodified reference RDA_solver (process_num=1: in-process ECOS, rda_solver.py:795-826) on
the instances of the committed oracle fixtures (tests/golden/oracle_*.npz: same seeds, same sizes, same
iteration counts), prints the trajectory gap reference <-> oracle fixture, and the ratio z / max(stuff, 0)
ECOS returns for inactive cells (the tie-break the oracle fixes at 1/2).  With --write it
stores the reference trajectories as tests/golden/true_reference_*.npz, which would turn the fixtures from
"oracle" into "reference" goldens.  Without cvxpy it prints one JSON line {"unavailable": ...} and exits 0.
"""
import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

FIXTURES = [   # (file, make_instance kwargs, solver kwargs, car kwargs) — as tests/golden/make_oracle_fixture*.py
    ('oracle_metric_T30N20.npz', dict(T=30, N=20, E=4), dict(), dict()),
    ('oracle_circles_T30N20.npz', dict(T=30, N=20, E=4, kind='circle', moving=True, lateral=(1.0, 6.0)),
     dict(min_sd=0.5, wu=0.2), dict(max_acce=(10, 1.0))),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    ap.add_argument('--write', action='store_true')
    args = ap.parse_args()
    try:
        import cvxpy  # noqa: F401
        import ecos  # noqa: F401
    except ImportError as ex:
        print(json.dumps({'unavailable': f'{ex.name} is not installed: the true reference cannot run here'}))
        return 0
    if not os.path.isdir(os.path.join(args.ref, 'RDA_planner')):
        print(json.dumps({'unavailable': f'no reference checkout at {args.ref}'}))
        return 0
    sys.path.insert(0, ROOT)
    from rda_planner_b200.scenarios import rectangle_robot, make_instance
    sys.path.insert(0, args.ref)
    for k in [m for m in sys.modules if m == 'RDA_planner' or m.startswith('RDA_planner.')]:
        del sys.modules[k]
    from RDA_planner.rda_solver import RDA_solver as RefSolver
    report = []
    for fname, inst_kw, solver_kw, car_kw in FIXTURES:
        fx = np.load(os.path.join(ROOT, 'tests', 'golden', fname))
        T, N = inst_kw['T'], inst_kw['N']
        car = rectangle_robot(**car_kw)
        inst = make_instance(int(fx['seed']), **inst_kw)
        ref_states = [inst['ref'][:, t:t + 1] for t in range(T + 1)]
        solver = RefSolver(T, car, max_edge_num=inst_kw['E'], max_obs_num=N, iter_num=int(fx['iters']), step_time=0.1,
                           iter_threshold=0.0, process_num=1, accelerated=True, time_print=False, **solver_kw)
        u, info = solver.iterative_solve(inst['nom_s'], inst['nom_u'], ref_states, inst['ref_speed'], list(inst['obstacles']))
        s = np.hstack(info['opt_state_list'])
        rec = {'fixture': fname, 'max_abs_du': float(np.abs(u - fx['u']).max()), 'max_abs_ds': float(np.abs(s - fx['s']).max())}
        try:    # tie-break actually taken by ECOS: z relative to the positive part of the margin, inactive cells
            z = np.array([p.value for p in solver.para_z_list])                 # N x (1 x T)
            rec['z_nonzero_fraction'] = float((np.abs(z) > 1e-9).mean())
        except Exception as ex:                                                  # attribute names may differ by version
            rec['z_probe'] = repr(ex)[:120]
        report.append(rec)
        if args.write:
            np.savez(os.path.join(ROOT, 'tests', 'golden', fname.replace('oracle_', 'true_reference_')),
                     seed=fx['seed'], iters=fx['iters'], u=u, s=s)
    print(json.dumps({'reference_vs_oracle_fixtures': report}))
    return 0


if __name__ == '__main__':
    sys.exit(main())
