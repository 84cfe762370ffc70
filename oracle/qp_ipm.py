"""TEST INFRASTRUCTURE — generic dense convex-QP solver used only by the oracle.

    minimise   0.5 x'Px + q'x     subject to   Gx <= h,   Ax = b

Mehrotra predictor-corrector primal-dual interior point method, float64, dense
numpy linear algebra.  It stands in for the third-party conic solver the
reference calls (cvxpy 1.5.2 -> ECOS, /root/reference/RDA_planner/rda_solver.py:693)
which is not installable in this image.  Nothing under rda_planner_b200/ may
import this module.
"""
import numpy as np


def solve_qp(P, q, G, h, A=None, b=None, tol=1e-9, max_iter=80):
    """Return (x, info).  info['status'] is 'optimal', 'optimal_inaccurate' or 'max_iter'."""
    P = np.asarray(P, float)
    q = np.asarray(q, float)
    G = np.asarray(G, float)
    h = np.asarray(h, float)
    n = q.size
    m = h.size
    if A is None:
        A = np.zeros((0, n))
        b = np.zeros(0)
    A = np.asarray(A, float)
    b = np.asarray(b, float)
    p = b.size

    def kkt_solve(W, r1, r2):
        # [P + G'WG  A'][dx]   [r1]
        # [A          0][dy] = [r2]
        H = P + G.T @ (W[:, None] * G)
        if p == 0:
            return np.linalg.solve(H, r1), np.zeros(0)
        K = np.block([[H, A.T], [A, np.zeros((p, p))]])
        sol = np.linalg.solve(K, np.concatenate([r1, r2]))
        return sol[:n], sol[n:]

    # initial point: regularised least-squares start, slacks/multipliers at least 1
    x, y = kkt_solve(np.ones(m), -q + G.T @ h, b)
    s = np.maximum(h - G @ x, 1.0)
    lam = np.ones(m)
    scale = 1.0 + max(np.abs(q).max(initial=0.0), np.abs(h).max(initial=0.0))
    status = 'max_iter'
    best = (np.inf, x, y, s, lam)
    it = 0
    for it in range(max_iter):
        rd = P @ x + q + G.T @ lam + A.T @ y
        rp = G @ x + s - h
        re = A @ x - b
        mu = float(s @ lam) / max(m, 1)
        merit = max(np.abs(rd).max(initial=0.0), np.abs(rp).max(initial=0.0),
                    np.abs(re).max(initial=0.0), mu)
        if merit < best[0]:
            best = (merit, x, y, s, lam)
        if merit < tol * scale:
            status = 'optimal'
            break
        if mu < 1e-14 * scale:          # complementarity exhausted: nothing left to gain
            break
        W = lam / s

        def direction(rc):
            # rc: complementarity residual target  s*dlam + lam*ds = -rc
            r1 = -rd - G.T @ ((lam * rp - rc) / s)
            dx, dy = kkt_solve(W, r1, -re)
            ds = -rp - G @ dx
            dl = -(rc + lam * ds) / s
            return dx, dy, ds, dl

        def max_step(v, dv):
            neg = dv < 0
            if not neg.any():
                return 1.0
            return min(1.0, float((-v[neg] / dv[neg]).min()))

        try:
            dxa, dya, dsa, dla = direction(s * lam)
            a_aff = min(max_step(s, dsa), max_step(lam, dla))
            mu_aff = float((s + a_aff * dsa) @ (lam + a_aff * dla)) / max(m, 1)
            sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
            dx, dy, ds, dl = direction(s * lam + dsa * dla - sigma * mu)
        except np.linalg.LinAlgError:
            break
        a = min(1.0, 0.99 * min(max_step(s, ds), max_step(lam, dl)))
        x = x + a * dx
        y = y + a * dy
        s = s + a * ds
        lam = lam + a * dl
    merit, x, y, s, lam = best
    if status != 'optimal' and merit < max(1e3 * tol, 1e-6) * scale:   # same acceptance whatever the requested tol
        status = 'optimal_inaccurate'      # accepted like cvxpy's OPTIMAL_INACCURATE (:696)
    return x, {'status': status, 'iters': it, 'merit': merit, 'ineq_dual': lam, 'eq_dual': y, 'slack': s}
