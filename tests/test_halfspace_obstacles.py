"""Obstacles given as GENERAL half-space sets (A, b) — what the reference accepts through MPC(rda_obstacle=True),
/root/reference/RDA_planner/mpc.py:150-155, and solves in the original rows (rda_solver.py:389-421): rows in any order,
redundant rows, unbounded sets (walls, wedges).  rda_planner_b200.rda_solver.canonical_polygon_rows reduces them to the
closed counter-clockwise polygon the kernels' geometry needs; the checks compare with oracle/cell_generic.py, which works on
the ORIGINAL rows and knows nothing about vertices or the closing square."""
import importlib.util
import os

import numpy as np
import pytest

import shim
from oracle import cpu_port
from oracle.cell_generic import solve_cell_generic
from rda_planner_b200.mpc import polygon_halfspaces
from rda_planner_b200.rda_solver import canonical_polygon_rows, pack_obstacles
from rda_planner_b200.scenarios import rectangle_robot

HERE = os.path.dirname(os.path.abspath(__file__))


def _gen():
    spec = importlib.util.spec_from_file_location('make_hs', os.path.join(HERE, 'golden', 'make_oracle_fixture_halfspace.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_canonical_rows_closed_polygon_is_untouched_and_general_sets_are_closed():
    V = np.array([[0, 2, 2, 0], [0, 0, 1, 1.0]])
    A, b = polygon_halfspaces(V)
    A2, b2 = canonical_polygon_rows(A, b)
    assert A2 is not None and np.array_equal(A2, np.asarray(A, float)) and np.array_equal(b2, np.ravel(b))
    # shuffled rows + a redundant row: the four supporting rows come back counter-clockwise, scaling kept
    perm = [2, 0, 3, 1]
    Ar = np.vstack([A[perm], [[1.0, 1.0]]])
    br = np.concatenate([np.ravel(b)[perm], [100.0]])
    A3, b3 = canonical_polygon_rows(Ar, br)
    assert A3.shape == (4, 2)
    ang = np.arctan2(A3[:, 1], A3[:, 0])
    assert np.all(np.mod(np.diff(np.concatenate([ang, ang[:1]])), 2 * np.pi) < np.pi)
    for a_, b_ in zip(A3, b3):
        assert any(np.allclose(a_, A[i]) and abs(b_ - np.ravel(b)[i]) < 1e-12 for i in range(4))
    # wall y <= -5 seen from a robot at (30, 0): closed by three sides of the 100 m square around the robot
    A4, b4 = canonical_polygon_rows(np.array([[0.0, 2.0]]), np.array([-10.0]), center=(30.0, 0.0))
    assert A4.shape == (4, 2) and any(np.allclose(a_, [0, 2.0]) and abs(b_ + 10.0) < 1e-12 for a_, b_ in zip(A4, b4))
    Vx = [np.linalg.solve(np.array([A4[i - 1], A4[i]]), np.array([b4[i - 1], b4[i]])) for i in range(4)]
    assert max(abs(v[0] - 30.0) for v in Vx) == pytest.approx(100.0) and min(v[1] for v in Vx) == pytest.approx(-100.0)
    # a set that does not reach into the default square gets a larger one instead of an error
    A5, b5 = canonical_polygon_rows(np.array([[1.0, 0.0]]), np.array([-300.0]))
    assert A5.shape[0] == 4
    with pytest.raises(ValueError):
        canonical_polygon_rows(np.array([[1.0, 0.0], [-1.0, 0.0]]), np.array([-1.0, -1.0]))      # x <= -1 and x >= 1
    # pack_obstacles reports the row budget an unbounded set needs
    from rda_planner_b200.mpc import rdaobs
    with pytest.raises(ValueError, match='max_edge_num'):
        pack_obstacles([rdaobs(np.array([[0.0, 1.0]]), np.array([[-5.0]]), 'Rpositive', None, None)], 4, 1, 3)


def test_cells_of_unbounded_sets_match_the_generic_oracle_on_the_original_rows():
    car = rectangle_robot()
    rng = np.random.default_rng(3)
    worst = {'d': 0.0, 'f': 0.0, 'lean': 0.0}
    lean_hits = 0
    for k in range(24):
        if k % 2 == 0:   # wall
            ang = rng.uniform(-3, 3)
            n = np.array([np.cos(ang), np.sin(ang)]) * rng.uniform(0.5, 3)
            A0, b0 = n[None], np.array([n @ rng.uniform(20, 60, 2)])
        else:            # wedge
            a1 = rng.uniform(-3, 3)
            a2 = a1 + rng.uniform(0.5, 2.5)
            A0 = np.array([[np.cos(a1), np.sin(a1)], [np.cos(a2), np.sin(a2)]])
            c = rng.uniform(20, 60, 2)
            b0 = A0 @ c
        A0 = A0.astype(np.float32).astype(float)
        b0 = b0.astype(np.float32).astype(float)
        n0 = A0[0] / np.linalg.norm(A0[0])
        foot = c if k % 2 else A0[0] * (b0[0] / (A0[0] @ A0[0]))
        p = foot + n0 * rng.uniform(0.5, 6.0) + np.array([-n0[1], n0[0]]) * rng.uniform(-3, 3) * (k % 2 == 0)
        Ac, bc = canonical_polygon_rows(A0, b0, center=p)
        phi, dbar = rng.uniform(-3, 3), rng.uniform(0.1, 1.0)
        zeta, xi = rng.uniform(-0.3, 0.3) * (k % 3 != 0), rng.uniform(-0.2, 0.2, 2) * (k % 4 == 1)
        ref = solve_cell_generic(A0, b0, False, car.G, car.h, p, phi, dbar, zeta, xi, 1.0)
        E = Ac.shape[0]
        lam_full = np.zeros(E)
        for i in range(A0.shape[0]):
            j = [q for q in range(E) if np.allclose(Ac[q], A0[i]) and abs(bc[q] - b0[i]) < 1e-6]
            assert j, 'an original row that supports the set was dropped'
            lam_full[j[0]] = ref['lam'][i]
        for prec in 'df':
            kk = shim.cell(car.G, car.h, 0, Ac, bc, p, phi, dbar, zeta, xi, 1.0, prec=prec)
            assert kk['path'] != 5
            err = max(np.abs(kk['lam'] - lam_full).max(), np.abs(kk['mu'] - ref['mu']).max(), abs(kk['z'] - ref['z']))
            worst[prec] = max(worst[prec], err)
        if E == 4:
            # the first pass of the GPU pipeline (cell_lean.cuh, float32) on the same rows
            out = np.zeros(28)
            fn = shim.lib().shim_cell_lean4
            fn.restype = shim.C.c_int
            G32, h32 = shim._f32(car.G), shim._f32(np.ravel(car.h))
            A32, b32 = shim._f32(Ac), shim._f32(bc)
            rc = fn(shim._p(G32), shim._p(h32), 4, 0, 4, shim._p(A32), shim._p(b32), shim.C.c_double(p[0]), shim.C.c_double(p[1]),
                    shim.C.c_double(phi), shim.C.c_double(dbar), shim.C.c_double(zeta), shim.C.c_double(xi[0]),
                    shim.C.c_double(xi[1]), shim.C.c_double(1.0), shim.C.c_double(0.5), shim._p(out))
            assert rc == 0
            if int(out[27]) == 0:
                lean_hits += 1
                err = max(np.abs(out[:4] - lam_full).max(), np.abs(out[8:12] - ref['mu']).max(), abs(out[16] - ref['z']))
                worst['lean'] = max(worst['lean'], err)
    # float32: a vertex of the closing square is up to 140 m from the robot, one ulp of direction is ~1e-5 m of margin there
    assert worst['d'] < 5e-6 and worst['f'] < 6e-4 and worst['lean'] < 6e-4, worst
    assert lean_hits >= 4


@pytest.mark.parametrize('name', ['w', 'x'])
def test_pipeline_with_walls_wedge_and_shuffled_rows_matches_the_oracle(name):
    """Whole ADMM loop: pack_obstacles (closing square around the robot) + compiled port of the kernel cores against the
    committed OracleRDA trace that used the ORIGINAL rows (tests/golden/make_oracle_fixture_halfspace.py)."""
    m = _gen()
    z = np.load(os.path.join(HERE, 'golden', 'oracle_halfspace.npz'))
    car, inst, T, N, iters = m.instance(name)
    E = 6
    A, b, kd, count, tv = pack_obstacles(list(inst['obstacles']), T, N, E, center=inst['nom_s'][0:2, 0], bound=30.0)
    for it in range(1, iters + 1):
        r = cpu_port.solve_batch(car, T, N, E, inst['nom_s'][None], inst['nom_u'][None], inst['ref'][None], [inst['ref_speed']],
                                 A[None], b[None], kd[None], [count], time_varying=tv, iter_num=it)
        assert r['cell_failures'][0, 0] == 0
        k = it - 1
        # float32 cells with vertices of the closing square 30-40 m away: ~1e-4 per cell, amplified over the iterations
        assert np.abs(r['s'][0] - z[f'{name}_s'][k]).max() < 1.5e-3
        assert np.abs(r['u'][0] - z[f'{name}_u'][k]).max() < 2e-3
        assert abs(r['resi_pri'][0] - z[f'{name}_resi_pri'][k]) < 2e-3 * (1 + z[f'{name}_resi_pri'][k])
    # resi_dual depends on the row scaling of lam: the walls keep their own rows, so it is comparable too
    assert abs(r['resi_dual'][0] - z[f'{name}_resi_dual'][-1]) < 5e-3 * (1 + z[f'{name}_resi_dual'][-1])


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['w', 'x'])
def test_gpu_with_walls_wedge_and_shuffled_rows_matches_the_oracle(name):
    from rda_planner_b200.rda_solver import RDA_solver
    m = _gen()
    z = np.load(os.path.join(HERE, 'golden', 'oracle_halfspace.npz'))
    car, inst, T, N, iters = m.instance(name)
    ref = [inst['ref'][:, t:t + 1] for t in range(T + 1)]
    for graph in (False, True):
        g = RDA_solver(T, car, max_edge_num=6, max_obs_num=N, iter_num=iters, iter_threshold=0.0, time_print=False, graph=graph)
        u, info = g.iterative_solve(inst['nom_s'], inst['nom_u'], ref, inst['ref_speed'], list(inst['obstacles']))
        assert info['status'] & 6 == 0
        assert np.abs(np.hstack(info['opt_state_list']) - z[f'{name}_s'][-1]).max() < 1.5e-3
        assert np.abs(u - z[f'{name}_u'][-1]).max() < 2e-3
        assert abs(info['resi_pri'] - z[f'{name}_resi_pri'][-1]) < 2e-3 * (1 + z[f'{name}_resi_pri'][-1])
        assert abs(info['resi_dual'] - z[f'{name}_resi_dual'][-1]) < 5e-3 * (1 + z[f'{name}_resi_dual'][-1])
