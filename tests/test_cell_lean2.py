"""Coherent first pass (cell_lean2.cuh, prepared for round 2) on the CPU build: with the support-vertex
pair of the previous pose it must either decline or reproduce the full lean pass / the float64 cell solver
at the new pose; a wrong feature pair must never be accepted with a wrong answer (separating-slab
certificate)."""
import numpy as np

import shim
from rda_planner_b200.scenarios import rectangle_robot
from rda_planner_b200.mpc import polygon_halfspaces
from rda_planner_b200.rda_solver import canonical_polygon_rows

KEYS = ('z', 'zeta_new', 'ax', 'ay', 'c0', 'gx', 'gy')


def _cases(seed, count):
    car = rectangle_robot()
    G, h = canonical_polygon_rows(car.G, car.h)
    rng = np.random.default_rng(seed)
    made = 0
    while made < count:
        n = int(rng.integers(3, 5))
        c = rng.uniform(20, 50, 2)
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        if np.min(np.diff(np.concatenate([ang, [ang[0] + 2 * np.pi]]))) < 0.3:
            continue
        V = c[:, None] + rng.uniform(0.5, 2.5) * np.vstack([np.cos(ang), np.sin(ang)])
        A, b = canonical_polygon_rows(*polygon_halfspaces(V))
        A4 = np.zeros((4, 2)); b4 = np.zeros(4); A4[:n] = A; b4[:n] = b
        a = rng.uniform(0, 2 * np.pi)
        p = c + rng.uniform(3, 9) * np.array([np.cos(a), np.sin(a)])
        phi = rng.uniform(-3.1, 3.1)
        dbar, zeta = rng.uniform(0.1, 1.0), rng.uniform(-0.2, 0.2)
        r0 = shim.cell(G, h, 0, A4, b4, p, phi, dbar, zeta, (0, 0), 1.0, prec='lean4')
        feat = int(r0['hm0'])
        if r0['path'] != 0 or not feat & 0x40:
            continue
        scale = rng.choice([0.01, 0.05, 0.3, 1.5])
        p1, phi1 = p + rng.normal(0, scale, 2), phi + rng.normal(0, 0.3 * scale)
        made += 1
        yield G, h, A4, b4, n, feat, p1, phi1, dbar, zeta, rng


def test_coherent_pass_matches_the_search_or_declines():
    hit = miss = 0
    for G, h, A4, b4, n, feat, p1, phi1, dbar, zeta, rng in _cases(5, 600):
        truth = shim.cell(G, h, 0, A4, b4, p1, phi1, dbar, zeta, (0, 0), 1.0, prec='d')
        full = shim.cell(G, h, 0, A4, b4, p1, phi1, dbar, zeta, (0, 0), 1.0, prec='lean4')
        r = shim.cell_lean2(G, h, A4, b4, feat, p1, phi1, dbar, zeta)
        if r['path'] == 6:
            miss += 1
            continue
        hit += 1
        assert r['feat'] & 0x40 and truth['path'] == 0
        assert max(abs(r[k] - truth[k]) for k in KEYS) < 2e-5
        np.testing.assert_allclose(r['lam'], truth['lam'][:4], atol=2e-5 * (1 + np.abs(truth['lam']).max()))
        np.testing.assert_allclose(r['mu'], truth['mu'][:4], atol=2e-5 * (1 + np.abs(truth['mu']).max()))
        if full['path'] == 0:
            assert max(abs(r[k] - full[k]) for k in KEYS) < 6e-5
    assert hit > 5 * miss            # coherence pays: > 80 % of the perturbed poses keep their feature pair


def test_wrong_feature_pairs_are_never_accepted_with_a_wrong_answer():
    accepted = 0
    for G, h, A4, b4, n, feat, p1, phi1, dbar, zeta, rng in _cases(6, 600):
        wrong = 0x40 | (int(rng.integers(0, n)) << 3) | int(rng.integers(0, 4))
        r = shim.cell_lean2(G, h, A4, b4, wrong, p1, phi1, dbar, zeta)
        if r['path'] == 6:
            continue
        accepted += 1
        truth = shim.cell(G, h, 0, A4, b4, p1, phi1, dbar, zeta, (0, 0), 1.0, prec='d')
        assert max(abs(r[k] - truth[k]) for k in KEYS) < 5e-5
    assert accepted > 0
    for bad in (0, 0x3f, 0x40 | (7 << 3) | 1, 0x40 | (1 << 3) | 7):      # invalid bytes decline
        G, h, A4, b4, n, feat, p1, phi1, dbar, zeta, rng = next(_cases(7, 1))
        if (bad >> 3) & 7 < n and bad & 7 < 4 and bad & 0x40:
            continue
        assert shim.cell_lean2(G, h, A4, b4, bad, p1, phi1, dbar, zeta)['path'] == 6
