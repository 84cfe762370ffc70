"""TEST INFRASTRUCTURE — "what an interior point solver returns" for ONE (obstacle, stage) cell.

The reference hands every per-obstacle LamMuZ problem (rda_solver.py:389-421) to cvxpy -> ECOS
(:768, :800), a primal-dual interior point method.  Where the argmin is not unique (hinge inactive:
the whole face {Hm + xi = 0, Im >= 0} is optimal) such a method does not return the
max-margin / LP-vertex point the CUDA kernels and oracle/cell_geo.py use; it follows the central path
and ends near its limit, the analytic centre of the optimal face.  ECOS itself is not installable in
this image, so this module restates that behaviour with a plain primal log-barrier path-following
method on the LITERAL cell problem in the original variables (lam, mu, z) plus the hinge epigraph
variable w >= neg(Im) cvxpy introduces for cp.neg:

    minimise   1/2 w^2 + ro2/2 |Hm|^2
               - tau [ sum log lam_i + sum log mu_j + log z + log w + log(w + Im) + log(1 - |A'lam|^2) ]

for tau -> 0 (same central path as a primal-dual method on the same constraint set; the exact ECOS
path additionally depends on cvxpy's canonicalisation — epigraph chains of cp.max / cp.norm shared by
the T stages of an obstacle — which is why this is an APPROXIMATION of the reference's tie-break,
used to bound its effect).  Polygon obstacles and polygon robots (Rpositive cones) only.
Only tests/ and tools/ may import this.
"""
import numpy as np


def solve_cell_ac(A, b, is_circle, G, h, p, phi, dbar, zeta, xi, ro2, theta=None, tau_end=1e-10):
    """Return dict(lam, mu, z, stuff, Hm, v, g, obj).  Arguments as oracle/cell_generic.solve_cell_generic."""
    if is_circle:
        raise NotImplementedError('cell_ac: polygon obstacles only')
    A = np.asarray(A, float)
    b = np.asarray(b, float).ravel()
    G = np.asarray(G, float)
    h = np.asarray(h, float).ravel()
    E, R = A.shape[0], G.shape[0]
    nrm = np.linalg.norm(A, axis=1)
    live = np.nonzero(nrm > 0)[0]
    ne = live.size
    An = A[live] / nrm[live, None]
    bn = (b[live] - A[live] @ p) / nrm[live]
    c, s_ = np.cos(phi), np.sin(phi)
    Rm = np.array([[c, -s_], [s_, c]])
    k0 = dbar - zeta
    n = ne + R + 2                                   # lam', mu, z, w
    iz, iw = ne + R, ne + R + 1
    M = np.zeros((2, n)); M[:, :ne] = (An @ Rm).T; M[:, ne:ne + R] = G.T          # Hm = M x + xi
    V = np.zeros((2, n)); V[:, :ne] = An.T                                          # v = V x
    aim = np.zeros(n); aim[:ne] = -bn; aim[ne:ne + R] = -h; aim[iz] = -1.0          # Im = aim.x - k0
    # linear positivity constraints  C x + c0 > 0 : lam', mu, z, w, w + Im
    C = np.vstack([np.eye(n), aim + np.eye(n)[iw]])
    c0 = np.concatenate([np.zeros(n), [-k0]])
    x = np.concatenate([np.full(ne, 0.1 / ne), np.full(R, 0.1), [1.0], [1.0]])
    Im0 = aim @ x - k0
    x[iw] = max(1.0, 1.0 - Im0)
    ew = np.eye(n)[iw]

    def fval(x, tau):
        lin = C @ x + c0
        v = V @ x
        q = 1.0 - v @ v
        if lin.min() <= 0 or q <= 0:
            return np.inf
        hm = M @ x + xi
        return 0.5 * x[iw] ** 2 + 0.5 * ro2 * hm @ hm - tau * (np.log(lin).sum() + np.log(q))

    tau = 1.0
    while True:
        for _ in range(60):
            lin = C @ x + c0
            v = V @ x
            q = 1.0 - v @ v
            hm = M @ x + xi
            g = x[iw] * ew + ro2 * M.T @ hm - tau * (C.T @ (1.0 / lin)) + tau * 2.0 * (V.T @ v) / q
            Vv = V.T @ v
            H = (np.outer(ew, ew) + ro2 * M.T @ M + tau * (C.T * (1.0 / lin ** 2)) @ C
                 + tau * (2.0 * V.T @ V / q + 4.0 * np.outer(Vv, Vv) / q ** 2))
            dx = np.linalg.solve(H + 1e-14 * np.eye(n), -g)
            dec = -g @ dx
            if dec < 1e-18 * max(1.0, tau) or dec < 1e-12 * tau:
                break
            t = 1.0
            f0 = fval(x, tau)
            while True:
                f1 = fval(x + t * dx, tau)
                if f1 <= f0 - 0.25 * t * dec or t < 1e-12:
                    break
                t *= 0.5
            x = x + t * dx
        if tau <= tau_end:
            break
        tau *= 0.2
    lamp = x[:ne]
    lam = np.zeros(E); lam[live] = lamp / nrm[live]
    mu = x[ne:ne + R].copy()
    z = float(x[iz])
    Im = aim @ x - k0
    hm = M @ x + xi
    return {'lam': lam, 'mu': mu, 'z': z, 'stuff': float(Im + z), 'Hm': hm - xi, 'v': V @ x, 'g': G.T @ mu,
            'obj': 0.5 * min(Im, 0.0) ** 2 + 0.5 * ro2 * float(hm @ hm), 'active': bool(Im < -1e-7)}
