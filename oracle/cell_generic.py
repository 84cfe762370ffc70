"""TEST INFRASTRUCTURE — generic solve of ONE (obstacle, stage) cell of the LamMuZ problem.

Restates /root/reference/RDA_planner/rda_solver.py:389-421 (LamMuZ_cost_cons),
:874-909 (Hm_LamMu, Im_LamMu) and :1034-1050 (cones).  The reference problem of one
obstacle is separable over the horizon (the max-of-norms constraint :408-416 is
equivalent to one norm constraint per stage), so the oracle solves it cell by cell
in the ORIGINAL variables (lam, mu, z) with scipy's SLSQP — no geometric reduction
is used here, which makes this file an independent check of the reduced form the
CUDA kernels use.

PARITY UNPINNED: the reference calls cvxpy/ECOS (absent in this image) and its
argmin is not unique whenever the hinge is inactive.  The documented tie-break is

  (L1) minimise the reference objective;
  (L2) among minimisers, maximise the margin  lam'(A p - b) - mu'h  (with Hm + xi = 0);
  (L3) represent (Aᵀlam, Gᵀmu) by the LP-vertex multipliers (min b'lam, min h'mu);
  (Z)  z = theta * max(stuff, 0), theta = 0.5 (analytic centre of [0, stuff]).
"""
import numpy as np
from scipy.optimize import minimize, linprog


def _row_scale(A):
    n = np.linalg.norm(A, axis=1)
    n = np.where(n > 0, n, 1.0)
    return n


def lp_vertex_poly(A, b, v):
    """min b'lam  s.t. A'lam = v, lam >= 0  (support-function multipliers)."""
    E = A.shape[0]
    if np.linalg.norm(v) < 1e-14:
        return np.zeros(E)
    sc = _row_scale(A)
    An = A / sc[:, None]
    bn = b / sc
    live = np.linalg.norm(A, axis=1) > 0
    bounds = [(0, None) if live[i] else (0, 0) for i in range(E)]
    res = linprog(bn, A_eq=An.T, b_eq=v, bounds=bounds, method='highs-ds')
    if res.status != 0:
        raise RuntimeError('lp_vertex_poly failed: %s' % res.message)
    return np.maximum(res.x, 0.0) / sc


def lam_from_v(A, b, is_circle, v):
    if is_circle:
        lam = np.zeros(A.shape[0])
        lam[0:2] = v
        lam[2] = -np.linalg.norm(v)
        return lam
    return lp_vertex_poly(A, b, v)


def mu_from_g(G, h, g, robot_cone='Rpositive'):
    """(L3) for the robot multipliers.  norm2 robot (rda_solver.py:1034-1039 with the ir-sim disc
    G = [[1,0],[0,1],[0,0]], h = (cx, cy, -r)): G'mu = mu[0:2] = g and -mu in the cone means
    |mu[0:2]| <= -mu[2]; min h'mu = g.c - r mu[2] is attained at mu[2] = -|g|."""
    if robot_cone == 'norm2':
        return np.array([g[0], g[1], -np.linalg.norm(g)])
    return lp_vertex_poly(G, h, g)


def solve_cell_generic(A, b, is_circle, G, h, p, phi, dbar, zeta, xi, ro2,
                       theta=0.5, robot_cone='Rpositive'):
    """Return dict(lam, mu, z, stuff, Hm, active).

    A (E,2), b (E,) obstacle copy t+1 (zero padded rows allowed); G (R,2), h (R,)
    polygon robot (Rpositive cone); p = nominal position s[0:2, t+1]; phi = nominal
    heading s[2, t] (NB column t, rda_solver.py:457-460, :555-562); dbar = d_t;
    zeta, xi(2,) current multipliers.
    """
    A = np.asarray(A, float)
    b = np.asarray(b, float).ravel()
    G = np.asarray(G, float)
    h = np.asarray(h, float).ravel()
    E, R = A.shape[0], G.shape[0]
    c, s_ = np.cos(phi), np.sin(phi)
    Rm = np.array([[c, -s_], [s_, c]])
    # shift the origin to the robot reference point: A p - b = -(b - A p)
    sc = _row_scale(A)
    An = A / sc[:, None]
    bn = (b - A @ p) / sc                     # b relative to p, row-normalised
    AR = An @ Rm
    k0 = dbar - zeta

    def split(th):
        return th[:E], th[E:]

    def margin(th):
        lam, mu = split(th)
        return -lam @ bn - mu @ h

    def hm(th):
        lam, mu = split(th)
        return G.T @ mu + AR.T @ lam + xi

    bounds = []
    for i in range(E):
        if is_circle:
            bounds.append((None, None) if i < 3 else (0, 0))
        else:
            bounds.append((0, None) if np.linalg.norm(A[i]) > 0 else (0, 0))
    disc_robot = robot_cone == 'norm2'
    if disc_robot:
        if R != 3 or np.abs(G - np.array([[1.0, 0], [0, 1], [0, 0]])).max() > 1e-12:
            raise NotImplementedError('norm2 robot: G must be [[1,0],[0,1],[0,0]] (ir-sim disc)')
        bounds += [(None, None)] * 3
    else:
        bounds += [(0, None)] * R
    cons = [{'type': 'ineq', 'fun': lambda th: 1.0 - np.sum((An.T @ th[:E]) ** 2),
             'jac': lambda th: np.concatenate([-2 * An @ (An.T @ th[:E]), np.zeros(R)])}]
    if is_circle:
        # -lam in norm2 cone: ||lam[0:2]|| + lam[2] <= 0   (rda_solver.py:1042-1050)
        cons.append({'type': 'ineq',
                     'fun': lambda th: -th[2] - np.sqrt(th[0] ** 2 + th[1] ** 2 + 1e-300)})
    if disc_robot:
        # -mu in norm2 cone: ||mu[0:2]|| + mu[2] <= 0   (rda_solver.py:1034-1039)
        cons.append({'type': 'ineq',
                     'fun': lambda th: -th[E + 2] - np.sqrt(th[E] ** 2 + th[E + 1] ** 2 + 1e-300)})
    # ---------- stage A: max margin with Hm + xi = 0 ----------
    consA = cons + [{'type': 'eq', 'fun': hm,
                     'jac': lambda th: np.hstack([AR.T, G.T])}]
    best, fallback = None, None
    for trial in range(8):
        rng = np.random.default_rng(trial)
        th0 = np.zeros(E + R) if trial == 0 else (0.1 if trial < 4 else 0.5) * rng.random(E + R)
        if is_circle:
            th0[2] = -1.0
            th0[3:E] = 0
        if disc_robot:
            th0[E + 2] = -1.0
        res = minimize(lambda th: -margin(th), th0, jac=lambda th: np.concatenate([bn, h]),
                       bounds=bounds, constraints=consA, method='SLSQP',
                       options={'ftol': 1e-14, 'maxiter': 400})
        viol = max(np.abs(hm(res.x)).max(), np.sum((An.T @ res.x[:E]) ** 2) - 1.0)
        if disc_robot:
            viol = max(viol, res.x[E + 2] + np.hypot(res.x[E], res.x[E + 1]))
        if viol < 1e-7 and (best is None or -res.fun > best[0]):
            best = (-res.fun, res.x)
        if fallback is None or viol < fallback[2]:
            fallback = (-res.fun, res.x, viol)
        if best is not None and trial >= 3:
            break
    if best is None:
        if fallback[2] < 1e-5:          # SLSQP stopped a hair outside the feasible set
            best = fallback[:2]
        else:
            raise RuntimeError('stage A failed (violation %.1e)' % fallback[2])
    cstar = best[0] - k0
    if cstar >= 0:
        th = best[1]
        v = An.T @ th[:E]
        g = G.T @ th[E:]
        active = False
        stuff = cstar
    else:
        # ---------- stage B: hinge active, z = 0 ----------
        def obj(th):
            m = margin(th) - k0
            q = hm(th)
            return 0.5 * min(m, 0.0) ** 2 + 0.5 * ro2 * q @ q

        def jac(th):
            m = margin(th) - k0
            q = hm(th)
            gm = np.concatenate([-bn, -h])
            return min(m, 0.0) * gm + ro2 * np.hstack([AR.T, G.T]).T @ q

        bestB = None
        starts = [best[1]] + [best[1] * (0.5 + 0.2 * k) for k in (1, 2, 3)]
        # the cone constraints are non-smooth at lam = 0 (overlapping sets end stage A there):
        # also start from small multipliers in a fan of directions
        for k in range(6):
            ang = np.pi * k / 3.0
            th0 = np.zeros(E + R)
            if is_circle:
                th0[0:3] = 0.05 * np.cos(ang), 0.05 * np.sin(ang), -0.05
            else:
                th0[:E] = 0.05 * np.maximum(An @ np.array([np.cos(ang), np.sin(ang)]), 0.0)
            th0[E:] = 0.01
            if disc_robot:
                th0[E:] = 0.01 * np.cos(ang + 1.0), 0.01 * np.sin(ang + 1.0), -0.02
            starts.append(th0)
        for th0 in starts:
            res = minimize(obj, th0, jac=jac, bounds=bounds, constraints=cons, method='SLSQP',
                           options={'ftol': 1e-15, 'maxiter': 600})
            if bestB is None or res.fun < bestB[0]:
                bestB = (res.fun, res.x)
        th = bestB[1]
        v = An.T @ th[:E]
        g = G.T @ th[E:]
        active = True
        stuff = None
    # ---------- (L3) LP-vertex multipliers ----------
    lam = lam_from_v(A, b, is_circle, v)
    mu = mu_from_g(G, h, g, robot_cone)
    m_val = lam @ (A @ p - b) - mu @ h - k0
    if active:
        z = 0.0
    else:
        z = theta * max(m_val, 0.0)
    Hm = G.T @ mu + (A @ Rm).T @ lam
    return {'lam': lam, 'mu': mu, 'z': z, 'stuff': m_val, 'Hm': Hm, 'active': active,
            'v': v, 'g': g}


def cell_objective(A, b, G, h, p, phi, dbar, zeta, xi, ro2, lam, mu, z):
    """Reference LamMuZ objective of one cell (accelerated mode), rda_solver.py:399-406."""
    c, s_ = np.cos(phi), np.sin(phi)
    Rm = np.array([[c, -s_], [s_, c]])
    Im = lam @ (A @ p) - lam @ b - mu @ h - dbar - z + zeta
    Hm = G.T @ mu + (A @ Rm).T @ lam + xi
    return 0.5 * min(Im, 0.0) ** 2 + 0.5 * ro2 * Hm @ Hm
