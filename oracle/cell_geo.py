"""TEST INFRASTRUCTURE — closed-form solve of the common cell cases (float64, numpy).

Used by the oracle to keep full-size runs tractable; every branch is cross-checked
against oracle/cell_generic.py (SLSQP in the original variables) by
tests/test_oracle_cell.py.  Cases solved here (all others -> generic solver):

  * xi == 0, obstacle and robot disjoint, hinge inactive (stuff >= 0): the
    max-margin certificate is the unit normal of the closest pair of points
    (polygon/polygon or disc/polygon distance), cf. SURVEY.md §9.7.
  * hinge active and the optimal robot contact point is a body VERTEX (KKT check).

Follows /root/reference/RDA_planner/rda_solver.py:389-421, 874-909.
"""
import numpy as np
from .cell_generic import solve_cell_generic, lam_from_v, lp_vertex_poly, mu_from_g


def poly_vertices(A, b):
    """Vertices of {x: Ax <= b} for CCW-ordered rows (row i = edge v_i -> v_{i+1},
    as produced by mpc.py:476-510).  Zero rows (padding) are ignored."""
    live = np.linalg.norm(A, axis=1) > 0
    A = A[live]
    b = np.asarray(b).ravel()[live]
    n = A.shape[0]
    V = np.zeros((n, 2))
    for i in range(n):
        M = np.array([A[i - 1], A[i]])
        V[i] = np.linalg.solve(M, np.array([b[i - 1], b[i]]))
    return V


def _closest_on_poly(P, V):
    """Closest point of polygon boundary/solid V (n,2 CCW) to point P; returns (x, inside)."""
    n = V.shape[0]
    best = (np.inf, None)
    inside = True
    for i in range(n):
        a, c = V[i], V[(i + 1) % n]
        e = c - a
        nrm = np.array([e[1], -e[0]])
        if nrm @ (P - a) > 0:
            inside = False
        t = np.clip((P - a) @ e / (e @ e), 0.0, 1.0)
        x = a + t * e
        d2 = (P - x) @ (P - x)
        if d2 < best[0]:
            best = (d2, x)
    return best[1], inside


def _separated(V1, V2):
    """SAT: True when some edge normal of either polygon strictly separates them."""
    for Va, Vb in ((V1, V2), (V2, V1)):
        n = Va.shape[0]
        for i in range(n):
            e = Va[(i + 1) % n] - Va[i]
            nrm = np.array([e[1], -e[0]])
            if np.min((Vb - Va[i]) @ nrm) > 1e-12 * (1 + np.linalg.norm(nrm)):
                return True
    return False


def _in_normal_cone(V, j, g, tol=1e-12):
    """g in the normal cone of CCW polygon V at vertex j (between outward normals of
    edges j-1 and j)."""
    n = V.shape[0]
    e_prev = V[j] - V[j - 1]
    e_next = V[(j + 1) % n] - V[j]
    # g.e_prev >= 0 and g.e_next <= 0  <=> V[j] maximises g.y locally (convex => globally)
    sc = np.linalg.norm(g) + 1e-300
    return (g @ e_prev >= -tol * sc * np.linalg.norm(e_prev)) and \
           (g @ e_next <= tol * sc * np.linalg.norm(e_next))


def solve_cell_geo(A, b, is_circle, G, h, p, phi, dbar, zeta, xi, ro2, theta=0.5,
                   Yb=None, stats=None):
    A = np.asarray(A, float)
    b = np.asarray(b, float).ravel()
    G = np.asarray(G, float)
    h = np.asarray(h, float).ravel()
    xi = np.asarray(xi, float).ravel()
    if Yb is None:
        Yb = poly_vertices(G, h)
    c, s_ = np.cos(phi), np.sin(phi)
    Rm = np.array([[c, -s_], [s_, c]])
    k0 = dbar - zeta
    Yw = p + Yb @ Rm.T
    Rn = Yb.shape[0]

    def finish(v, g, active, tag):
        lam = lam_from_v(A, b, is_circle, v)
        mu = lp_vertex_poly(G, h, g)
        stuff = lam @ (A @ p - b) - mu @ h - k0
        z = 0.0 if active else theta * max(stuff, 0.0)
        Hm = G.T @ mu + (A @ Rm).T @ lam
        if stats is not None:
            stats[tag] = stats.get(tag, 0) + 1
        return {'lam': lam, 'mu': mu, 'z': z, 'stuff': stuff, 'Hm': Hm, 'active': active,
                'v': v, 'g': g}

    # --- closest point of the obstacle to each robot vertex, and robot to obstacle ---
    if is_circle:
        ctr = b[0:2].copy()
        rad = -b[2]
        xrob, inside = _closest_on_poly(ctr, Yw)      # closest robot point to the centre
        dist_c = np.linalg.norm(xrob - ctr)
        disjoint = (not inside) and dist_c > rad
        if disjoint:
            vdir = (xrob - ctr) / dist_c
            dist = dist_c - rad
        def near_obs(Pt):
            dd = np.linalg.norm(Pt - ctr)
            return ctr + rad * (Pt - ctr) / dd, dd - rad
    else:
        Vo = poly_vertices(A, b)
        disjoint = _separated(Vo, Yw)
        if disjoint:
            best = (np.inf, None, None)
            for j in range(Rn):
                x, _ = _closest_on_poly(Yw[j], Vo)
                d2 = np.sum((Yw[j] - x) ** 2)
                if d2 < best[0]:
                    best = (d2, x, Yw[j])
            for i in range(Vo.shape[0]):
                y, _ = _closest_on_poly(Vo[i], Yw)
                d2 = np.sum((y - Vo[i]) ** 2)
                if d2 < best[0]:
                    best = (d2, Vo[i], y)
            dist = np.sqrt(best[0])
            vdir = (best[2] - best[1]) / dist
        def near_obs(Pt):
            x, ins = _closest_on_poly(Pt, Vo)
            return x, (0.0 if ins else np.linalg.norm(Pt - x))

    if disjoint and not np.any(xi != 0):
        cstar = dist - k0
        if cstar >= 0:
            return finish(vdir, -Rm.T @ vdir, False, 'geo_inactive')
    if disjoint:
        # vertex-contact candidates with KKT check (sufficient: the problem is convex)
        for j in range(Rn):
            x, dj = near_obs(Yw[j])
            if dj <= 1e-12:
                continue
            vj = (Yw[j] - x) / dj
            Dj = dj + xi @ Yb[j] - k0
            if Dj >= 0:
                g = -Rm.T @ vj - xi
                if _in_normal_cone(Yb, j, g):
                    # stage-A optimum with non-negative margin -> inactive
                    # (only valid as the max-margin point if it is stage-A optimal: KKT ok)
                    return finish(vj, g, False, 'geo_inactive_vertex')
            else:
                tau = -Dj / (1.0 + Yb[j] @ Yb[j] / ro2)
                q = -tau * Yb[j] / ro2
                g = q - Rm.T @ vj - xi
                if _in_normal_cone(Yb, j, g):
                    # must also be the stage-A decision "active": margin max < 0.  The
                    # stage-B KKT point with m < 0 is the global optimum of (L1) and its
                    # optimal value is > 0, hence no inactive point exists.
                    return finish(vj, g, True, 'geo_active_vertex')
    if stats is not None:
        stats['generic'] = stats.get('generic', 0) + 1
    return solve_cell_generic(A, b, is_circle, G, h, p, phi, dbar, zeta, xi, ro2, theta)


def solve_cell_geo_disc(A, b, is_circle, G, h, p, phi, dbar, zeta, xi, ro2, theta=0.5, stats=None):
    """DISC body (cone_type 'norm2', rda_solver.py:1034-1039; G = [[1,0],[0,1],[0,0]], h = (cx, cy, -r)).
    Closed form only for xi == 0, disjoint sets and an inactive hinge: the closest pair lies on the line from the
    closest obstacle point to the disc centre, margin = dist(centre, O) - r.  Everything else -> generic solver."""
    A = np.asarray(A, float)
    b = np.asarray(b, float).ravel()
    G = np.asarray(G, float)
    h = np.asarray(h, float).ravel()
    xi = np.asarray(xi, float).ravel()
    c, s_ = np.cos(phi), np.sin(phi)
    Rm = np.array([[c, -s_], [s_, c]])
    k0 = dbar - zeta
    ctr = p + Rm @ h[0:2]
    rr = -h[2]
    if not np.any(xi != 0):
        if is_circle:
            d = ctr - b[0:2]
            dn = np.linalg.norm(d)
            gap = dn - (-b[2]) - rr
            vdir = d / dn if dn > 0 else None
        else:
            x, inside = _closest_on_poly(ctr, poly_vertices(A, b))
            dn = np.linalg.norm(ctr - x)
            gap = -1.0 if inside else dn - rr
            vdir = (ctr - x) / dn if dn > 0 else None
        if gap > 1e-9 and gap - k0 >= 0 and vdir is not None:
            g = -Rm.T @ vdir
            lam = lam_from_v(A, b, is_circle, vdir)
            mu = mu_from_g(G, h, g, 'norm2')
            stuff = lam @ (A @ p - b) - mu @ h - k0
            if stats is not None:
                stats['geo_disc_inactive'] = stats.get('geo_disc_inactive', 0) + 1
            return {'lam': lam, 'mu': mu, 'z': theta * max(stuff, 0.0), 'stuff': stuff,
                    'Hm': G.T @ mu + (A @ Rm).T @ lam, 'active': False, 'v': vdir, 'g': g}
    if stats is not None:
        stats['generic'] = stats.get('generic', 0) + 1
    return solve_cell_generic(A, b, is_circle, G, h, p, phi, dbar, zeta, xi, ro2, theta, robot_cone='norm2')
