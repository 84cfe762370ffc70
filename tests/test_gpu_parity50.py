"""-m gpu: parity at the BENCHMARKED configuration.  24 metric-row instances (T=30, N=20 polygons,
E=4), 50 ADMM iterations, early stop disabled, against the per-iteration oracle trace tests/golden/oracle_metric50.npz
(16 instances of the bench workload + 8 of the harsher generator with obstacles on the path).  The test REPORTS the
distribution of the trajectory / control / residual / su-objective gaps per iteration (also written to
gpurun_out/parity50_r02.json) and bounds it:

  * up to 8 iterations every instance stays inside the stated float32 tolerance (states 1e-3; controls 5e-3: the su-QP
    is nearly flat in some control directions — the oracle's own two solvers differ there by 3e-4);
    (the trace was regenerated in the third session of round 2 with the oracle's su-QP solved to 1e-12 instead of 1e-9 — the
    oracle had been the less accurate side; CPU build of the same cores against the new trace,
    tests/test_parity50_cpu.py: states max 2.8e-4, controls max 7.8e-4 through 8 iterations; the bounds were left as they were);
  * beyond, the ADMM map is not contractive on a few instances (the float64 oracle itself moves by more
    than 1e-3 when its su-QP start is perturbed), so the bound is on the BULK of the distribution (median and the
    75th percentile) while the maximum is reported, not bounded.
"""
import json
import os

import numpy as np
import pytest
import torch

from rda_planner_b200.scenarios import rectangle_robot, make_instance

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
T, N, E = 30, 20, 4
CHECK = [1, 2, 4, 8, 16, 32, 50]


def literal_su_cost(s, u, d, lin_s, ref, vref, coef, pref, car_h, tun):
    """The reference su cost (rda_solver.py:313-387 with :831-872, :1011-1032) from the DEVICE buffers the su-QP
    kernel reads: coef [5][N][T] = (lam'A x, lam'A y, c0, (mu'G + xi) x, y), c0 relative to pref."""
    ws, wu, slack_gain, ro1, ro2 = tun
    nav = ws * np.sum((s - ref) ** 2) + wu * np.sum((u[0] - vref) ** 2) - slack_gain * d.sum()
    ax, ay, c0, gx, gy = coef
    dx = s[0, 1:][None] - pref[0][None]
    dy = s[1, 1:][None] - pref[1][None]
    Im = ax * dx + ay * dy + c0 - d[None]
    hinge = 0.5 * ro1 * np.sum(np.minimum(Im, 0.0) ** 2)
    ph = lin_s[2, :-1]
    c, sn = np.cos(ph)[None], np.sin(ph)[None]
    dphi = (s[2, 1:] - ph)[None]
    # a (R - phib R' + R' phi) = a R + (phi - phib) a R'
    h0 = gx + (ax * c + ay * sn) + dphi * (-ax * sn + ay * c)
    h1 = gy + (-ax * sn + ay * c) + dphi * (-ax * c - ay * sn)
    cons = 0.5 * ro2 * np.sum(h0 ** 2 + h1 ** 2)
    return nav + hinge + cons, nav, hinge, cons


def test_distribution_of_gaps_over_50_iterations():
    from rda_planner_b200.rda_solver import RDA_solver, pack_obstacles
    from rda_planner_b200 import _cabi
    z = np.load(os.path.join(HERE, 'golden', 'oracle_metric50.npz'))
    B = len(z['seeds'])
    car = rectangle_robot()
    insts = [make_instance(int(sd), T=T, N=N, E=E, lateral=tuple(l)) for sd, l in zip(z['seeds'], z['lateral'])]
    packs = [pack_obstacles(list(i['obstacles']), T, N, E) for i in insts]
    st = lambda k: np.stack([i[k] for i in insts]).astype(np.float32)
    g = RDA_solver(T, car, max_edge_num=E, max_obs_num=N, iter_num=50, iter_threshold=0.0, time_print=False, batch=B)
    g.begin(st('nom_s'), st('nom_u'), st('ref'), np.array([i['ref_speed'] for i in insts], np.float32),
            np.stack([p[0] for p in packs]), np.stack([p[1] for p in packs]), np.stack([p[2] for p in packs]),
            np.array([p[3] for p in packs], np.int32), False, 0.0)
    tun = (1.0, 1.0, 8.0, 200.0, 1.0)
    ref = st('ref').astype(float)
    rows = []
    for it in range(1, 51):
        if it in CHECK:
            coef = g.state_buffer(_cabi.BUF_COEF, (B, 5, N, T)).double().cpu().numpy()
            pref = g.state_buffer(_cabi.BUF_PREF, (B, 2, T)).double().cpu().numpy()
            lin_s = g.state_buffer(_cabi.BUF_CUR_S, (B, 3, T + 1)).double().cpu().numpy()
        g.step_su()
        if it in CHECK:
            s = g.state_buffer(_cabi.BUF_CUR_S, (B, 3, T + 1)).double().cpu().numpy()
            u = g.state_buffer(_cabi.BUF_CUR_U, (B, 2, T)).double().cpu().numpy()
            d = g.state_buffer(_cabi.BUF_DIS, (B, T)).double().cpu().numpy()
        g.step_lammuz()
        if it not in CHECK:
            continue
        out = {k: v.clone() for k, v in g.finish().items()}
        assert int((out['status'] & 6).sum()) == 0, 'an instance kept a previous iterate'
        k = it - 1
        ds = np.abs(s - z['s'][:, k]).reshape(B, -1).max(1)
        du = np.abs(u - z['u'][:, k]).reshape(B, -1).max(1)
        dd = np.abs(d - z['d'][:, k]).reshape(B, -1).max(1)
        cost = np.array([literal_su_cost(s[b], u[b], d[b], lin_s[b], ref[b], 4.0, coef[b], pref[b], None, tun) for b in range(B)])
        oc = z['cost'][:, k]
        dcost = np.abs(cost[:, 0] - oc[:, 0]) / (1.0 + np.abs(oc[:, 0]))
        dnav = np.abs(cost[:, 1] - oc[:, 1]) / (1.0 + np.abs(oc[:, 1]))
        rp = np.abs(out['resi_pri'].cpu().numpy() - z['resi_pri'][:, k]) / (1 + z['resi_pri'][:, k])
        rd = np.abs(out['resi_dual'].cpu().numpy() - z['resi_dual'][:, k]) / (1 + z['resi_dual'][:, k])
        active = oc[:, 2] > 1e-6          # instances whose su-QP ends with active hinges (contact / overlap)
        q = lambda x: {'median': float(np.median(x)), 'p75': float(np.quantile(x, .75)), 'p95': float(np.quantile(x, .95)),
                       'max': float(x.max())}
        rows.append({'iteration': it, 'state_gap': q(ds), 'control_gap': q(du), 'd_gap': q(dd), 'su_cost_rel_gap': q(dcost),
                     'nav_cost_rel_gap': q(dnav), 'resi_pri_rel_gap': q(rp), 'resi_dual_rel_gap': q(rd),
                     'state_gap_bench16': q(ds[:16]), 'state_gap_harsh8': q(ds[16:]),
                     'state_gap_active_hinge': q(ds[active]) if active.any() else None,
                     'state_gap_inactive': q(ds[~active]) if (~active).any() else None,
                     'instances_with_active_hinges': int(active.sum())})
        print(f"it {it:2d}: |ds| med {np.median(ds):.1e} p75 {np.quantile(ds, .75):.1e} max {ds.max():.1e}   |du| med {np.median(du):.1e} "
              f"max {du.max():.1e}   cost med {np.median(dcost):.1e} max {dcost.max():.1e}   resi_pri med {np.median(rp):.1e} "
              f"resi_dual med {np.median(rd):.1e}")
        if it <= 8:
            assert ds.max() < 1e-3 and du.max() < 5e-3 and dd.max() < 5e-3, (it, ds.max(), du.max(), dd.max())
            assert np.median(dcost) < 1e-4 and dcost.max() < 5e-3, (it, dcost.max())
            assert rp.max() < 2e-3 and rd.max() < 2e-3
        else:
            assert np.median(ds) < 2e-3 and np.quantile(ds, .75) < 2e-2, (it, np.median(ds), np.quantile(ds, .75))
            assert np.median(du) < 5e-3 and np.median(dcost) < 1e-3 and np.median(rp) < 2e-3 and np.median(rd) < 2e-3
        assert np.all(np.isfinite(s)) and np.all(np.isfinite(u))
    os.makedirs(os.path.join(HERE, '..', 'gpurun_out'), exist_ok=True)
    with open(os.path.join(HERE, '..', 'gpurun_out', 'parity50_r02.json'), 'w') as f:
        json.dump({'instances': B, 'fixture': 'tests/golden/oracle_metric50.npz', 'rows': rows}, f, indent=1)
