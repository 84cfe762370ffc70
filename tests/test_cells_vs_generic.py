"""The kernels' (lam, mu, z) cell core (csrc/cell_solver.cuh, g++ build) against the GENERIC solver of the cell in the
reference's ORIGINAL variables (oracle/cell_generic.py: SLSQP + HiGHS on rda_solver.py:389-421, no geometry) — not against
oracle/cell_geo.py, which shares the closed-form idea with the kernels.  Random cells of the three
regimes the metric workload produces (far, near / active hinge, overlapping sets), polygon (3-4 rows, zero padded) and disc
obstacles, random multipliers xi / zeta, three values of ro2; float64 and float32 arithmetic of the core."""
import numpy as np
import pytest

import shim
from rda_planner_b200.scenarios import rectangle_robot
from rda_planner_b200.mpc import polygon_halfspaces
from oracle.cell_generic import solve_cell_generic, cell_objective

# (lam, mu, z) gap, every cell / 99th percentile.  Measured (1 140 cells): float64 core max 5e-7; float32 core median 2e-7,
# p99 2.5e-6, max 1.7e-4 (one active cell whose contact is an obstacle vertex: float32 resolves the contact direction to ~1e-4).
TOL = {'d': 2e-6, 'f': 3e-4}
TOL_P99 = {'d': 1e-6, 'f': 2e-5}
OBJ_TOL = {'d': 1e-9, 'f': 1e-4}     # reference objective (rda_solver.py:399-406) of the core's point above the generic optimum


def _random_cell(rng, k, kind):
    cls = k % 3                                          # 0 far, 1 near, 2 overlapping
    dist = (rng.uniform(4, 12), rng.uniform(0.2, 3.0), rng.uniform(0.0, 1.5))[cls]
    th = rng.uniform(0, 2 * np.pi)
    base = np.array([1.5, 0.0]) + (dist + (2.5 if cls < 2 else 0.0)) * np.array([np.cos(th), np.sin(th)])
    phi = rng.uniform(-np.pi, np.pi)
    p = rng.uniform(-30, 30, 2)
    c, s = np.cos(phi), np.sin(phi)
    ctr = p + np.array([[c, -s], [s, c]]) @ base
    if kind == 'circle':
        rad = rng.uniform(0.3, 1.5)
        A = np.array([[1.0, 0], [0, 1.0], [0, 0], [0, 0]])
        b = np.array([[ctr[0]], [ctr[1]], [-rad], [0.0]])
    else:
        nv = int(rng.integers(3, 5))
        ang = np.sort(np.linspace(0, 2 * np.pi, nv, endpoint=False) + rng.uniform(-0.3, 0.3, nv) + rng.uniform(0, 6.28))
        r = rng.uniform(0.5, 2.0)
        V = np.stack([ctr[0] + r * np.cos(ang) * rng.uniform(0.7, 1.3), ctr[1] + r * np.sin(ang) * rng.uniform(0.7, 1.3)])
        A, b = polygon_halfspaces(V)
        A = np.vstack([A, np.zeros((4 - nv, 2))])
        b = np.vstack([b.reshape(-1, 1), np.zeros((4 - nv, 1))])
    A = A.astype(np.float32).astype(float)               # the rows both solvers see are the float32 the kernels store
    b = b.astype(np.float32).astype(float)
    dbar = rng.uniform(0.1, 1.0)
    zeta = rng.normal(0, 0.3) * (rng.random() < 0.7)
    xi = rng.normal(0, 0.2, 2) * (rng.random() < 0.5)
    ro2 = (0.5, 1.0, 5.0)[int(rng.integers(0, 3))]
    return A, b, p, phi, dbar, zeta, xi, ro2


@pytest.mark.parametrize('kind,seed,n', [('polygon', 101, 450), ('polygon', 102, 450), ('circle', 103, 60)])
def test_cell_core_equals_generic_solver_on_random_cells(kind, seed, n):
    car = rectangle_robot()
    G, h = car.G, np.asarray(car.h).ravel()
    rng = np.random.default_rng(seed)
    seen = {'active': 0, 'inactive': 0, 'tilted': 0, 'paths': set()}
    gaps = {'d': [], 'f': []}
    for k in range(n):
        A, b, p, phi, dbar, zeta, xi, ro2 = _random_cell(rng, k, kind)
        circ = kind == 'circle'
        r = solve_cell_generic(A, b, circ, G, h, p, phi, dbar, zeta, xi, ro2)
        fo = cell_objective(A, b.ravel(), G, h, p, phi, dbar, zeta, xi, ro2, r['lam'], r['mu'], r['z'])
        seen['active' if r['active'] else 'inactive'] += 1
        seen['tilted'] += bool(np.any(xi != 0))
        for prec in ('d', 'f'):
            kk = shim.cell(G, h, int(circ), A, b, p, phi, dbar, zeta, xi, ro2, prec=prec)
            assert kk['path'] != 5, (k, prec)                                   # 5: the core gave up (keep-previous)
            seen['paths'].add(kk['path'])
            fk = cell_objective(A, b.ravel(), G, h, p, phi, dbar, zeta, xi, ro2, kk['lam'], kk['mu'], kk['z'])
            assert fk - fo < OBJ_TOL[prec] * (1 + fo), (k, prec, fk, fo)
            np.testing.assert_allclose(kk['lam'], r['lam'], atol=TOL[prec], err_msg=f'cell {k} {prec}')
            np.testing.assert_allclose(kk['mu'], r['mu'], atol=TOL[prec], err_msg=f'cell {k} {prec}')
            assert abs(kk['z'] - r['z']) < TOL[prec], (k, prec)
            gaps[prec].append(max(np.abs(kk['lam'] - r['lam']).max(), np.abs(kk['mu'] - r['mu']).max()))
            # feasibility of the reference constraints (:408-419) at the core's point
            if circ:
                assert np.hypot(kk['lam'][0], kk['lam'][1]) <= -kk['lam'][2] + 1e-6
                assert np.hypot(kk['lam'][0], kk['lam'][1]) <= 1 + 1e-5
            else:
                assert (kk['lam'] >= -1e-7).all() and np.linalg.norm(A.T @ kk['lam']) <= 1 + 1e-5
            assert (kk['mu'] >= -1e-7).all() and kk['z'] >= 0
    for prec in ('d', 'f'):
        assert np.quantile(gaps[prec], 0.99) < TOL_P99[prec]
    assert seen['active'] > n // 5 and seen['inactive'] > n // 5 and seen['tilted'] > n // 4 and len(seen['paths']) >= 2
