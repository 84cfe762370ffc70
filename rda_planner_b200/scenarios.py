"""Seeded synthetic planning instances shared by bench.py and the tests.

Everything here is plain numpy on the host; it only produces the inputs of
RDA_solver.iterative_solve / iterative_solve_batch (nominal states and controls,
reference points, obstacle half-spaces)."""
from collections import namedtuple
from math import cos, sin, tan

import numpy as np

from .mpc import polygon_halfspaces, rdaobs

car = namedtuple('car', 'G h cone_type wheelbase max_speed max_acce dynamics')


def rectangle_robot(length=4.6, width=1.6, wheelbase=3.0, dynamics='acker',
                    max_speed=(10, 1), max_acce=(10, 0.5)):
    """Car tuple for a rectangular body whose reference point is the rear axle centre
    (ir-sim convention): x in [-(length-wheelbase)/2, (length+wheelbase)/2]."""
    x0 = -(length - wheelbase) / 2
    x1 = (length + wheelbase) / 2
    y1 = width / 2
    vert = np.array([[x0, x1, x1, x0], [-y1, -y1, y1, y1]])
    G, h = polygon_halfspaces(vert)
    return car(G, h, 'Rpositive', wheelbase, list(max_speed), list(max_acce), dynamics)


def disc_robot(radius=1.0, center=(0.0, 0.0), wheelbase=1.0, dynamics='diff', max_speed=(10, 1), max_acce=(10, 0.5)):
    """Car tuple of a circular body as ir-sim describes it (cone_type 'norm2', rda_solver.py:1034-1039):
    G = [[1,0],[0,1],[0,0]], h = (cx, cy, -r), i.e. |y - c| <= r in the body frame."""
    G = np.array([[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]])
    h = np.array([[center[0]], [center[1]], [-radius]])
    return car(G, h, 'norm2', wheelbase, list(max_speed), list(max_acce), dynamics)


def rollout(state, u, dt, L, dynamics):
    """Nominal trajectory of the nonlinear model (mpc.py:293-336) for controls u (2,T)."""
    T = u.shape[1]
    s = np.zeros((3, T + 1))
    s[:, 0] = np.asarray(state, float).reshape(3)
    for t in range(T):
        v, w = u[0, t], u[1, t]
        th = s[2, t]
        if dynamics == 'acker':
            ds = np.array([v * cos(th), v * sin(th), v * tan(w) / L])
        elif dynamics == 'diff':
            ds = np.array([v * cos(th), v * sin(th), w])
        else:
            ds = np.array([v * cos(w), v * sin(w), 0.0])
        s[:, t + 1] = s[:, t] + ds * dt
    return s


def rect_vertices(cx, cy, length, width, yaw):
    c, s = cos(yaw), sin(yaw)
    loc = np.array([[-length / 2, length / 2, length / 2, -length / 2],
                    [-width / 2, -width / 2, width / 2, width / 2]])
    return np.array([[c, -s], [s, c]]) @ loc + np.array([[cx], [cy]])


def random_convex_polygon(rng, cx, cy, radius, k):
    ang = np.sort(rng.uniform(0, 2 * np.pi, k))
    # keep it well-conditioned: spread the angles
    ang = np.linspace(0, 2 * np.pi, k, endpoint=False) + rng.uniform(-0.3, 0.3, k) * (2 * np.pi / k)
    return np.array([cx + radius * np.cos(ang), cy + radius * np.sin(ang)])


def make_instance(seed, T=30, N=20, E=4, dynamics='acker', dt=0.1, ref_speed=4.0,
                  kind='polygon', moving=False, lateral=(1.8, 6.0), v_nom=None):
    """One planning instance along a straight reference line (corridor geometry).

    Returns dict(nom_s (3,T+1), nom_u (2,T), ref (3,T+1), ref_speed, obstacles [rdaobs]*N).
    Obstacles are boxes / convex k-gons / discs scattered beside and on the path ahead of
    the robot so that a few of them constrain the motion."""
    rng = np.random.default_rng(seed)
    L = 3.0
    heading = rng.uniform(-np.pi, np.pi)
    start = np.array([rng.uniform(5, 55), rng.uniform(5, 55)])
    dirv = np.array([cos(heading), sin(heading)])
    nrm = np.array([-dirv[1], dirv[0]])
    state = np.array([start[0] + rng.normal(0, 0.3) * nrm[0], start[1] + rng.normal(0, 0.3) * nrm[1],
                      heading + rng.normal(0, 0.1)])
    v0 = ref_speed if v_nom is None else v_nom
    nom_u = np.vstack([np.full(T, v0), np.zeros(T)])
    if dynamics == 'omni':
        nom_u[1] = heading
    nom_s = rollout(state, nom_u, dt, L, dynamics)
    ref = np.zeros((3, T + 1))
    for t in range(T + 1):
        ref[0:2, t] = start + dirv * (ref_speed * dt * t)
        ref[2, t] = heading
    reach = ref_speed * dt * T
    obstacles = []
    for o in range(N):
        along = rng.uniform(2.0, reach + 6.0)
        side = rng.choice([-1.0, 1.0])
        lat = side * rng.uniform(*lateral)
        c = start + dirv * along + nrm * lat
        vel = np.zeros((2, 1))
        if moving:
            sp = rng.uniform(0, 1.0)
            hd = rng.uniform(-np.pi, np.pi)
            vel = np.array([[sp * cos(hd)], [sp * sin(hd)]])
        if kind == 'circle':
            r = rng.uniform(0.5, 1.0)
            A0 = np.array([[1.0, 0], [0, 1.0], [0, 0]])
            if moving and np.linalg.norm(vel) > 0.01:
                A = [A0.copy() for _ in range(T + 1)]
                b = [np.vstack((c.reshape(2, 1) + vel * (t * dt), [[-r]])) for t in range(T + 1)]
            else:
                A, b = A0, np.vstack((c.reshape(2, 1), [[-r]]))
            obstacles.append(rdaobs(A, b, 'norm2', c.reshape(2, 1), None))
        else:
            if E == 4:
                vert = rect_vertices(c[0], c[1], rng.uniform(1.5, 5.0), rng.uniform(1.0, 2.5),
                                     rng.uniform(0, np.pi))
            else:
                k = int(rng.integers(3, E + 1))
                vert = random_convex_polygon(rng, c[0], c[1], rng.uniform(0.5, 2.0), k)
            if moving and np.linalg.norm(vel) > 0.01:
                A, b = [], []
                for t in range(T + 1):
                    At, bt = polygon_halfspaces(vert + vel * (t * dt))
                    A.append(At); b.append(bt)
            else:
                A, b = polygon_halfspaces(vert)
            obstacles.append(rdaobs(A, b, 'Rpositive', None, vert))
    return {'nom_s': nom_s, 'nom_u': nom_u, 'ref': ref, 'ref_speed': ref_speed,
            'obstacles': obstacles, 'state': state}


# ---- BASELINE.json configs: seeded synthetic instances of the five named workloads --------------
# name -> solver shape, tunables, global batch and GPU count the config is quoted on
CONFIGS = {
    'A': dict(what='path_track: acker, T=10, 4 static polygons, 1 instance', T=10, N=4, E=4, dynamics='acker', iter_num=2,
              tun=dict(ro1=300), global_batch=1, gpus=1),
    'B': dict(what='corridor: diff-drive, T=20, 10 polygon obstacles (two 70 x 2 m walls + 8 boxes), batch 64', T=20, N=10, E=4,
              dynamics='diff', iter_num=50, tun={}, global_batch=64, gpus=1),
    'C': dict(what='dynamic_obs: acker, T=30, 20 moving discs (per-stage copies), batch 512', T=30, N=20, E=3, dynamics='acker',
              iter_num=50, tun=dict(min_sd=0.5, wu=0.2), global_batch=512, gpus=1),
    'D': dict(what='lidar: acker, T=30, 64 convex hulls (3-8 vertices) within 10 m, batch 2048 over 4 GPUs', T=30, N=64, E=8,
              dynamics='acker', iter_num=50, tun=dict(slack_gain=13), global_batch=2048, gpus=4),
    'E': dict(what='stress: acker, T=40, 128 convex polytopes (<= 8 faces) in a 40 x 40 m field, batch 8192 over 8 GPUs', T=40,
              N=128, E=8, dynamics='acker', iter_num=50, tun={}, global_batch=8192, gpus=8),
    'metric': dict(what='metric row: acker, T=30, N=20 static polygons (E=4)', T=30, N=20, E=4, dynamics='acker', iter_num=50,
                   tun={}, global_batch=None, gpus=None),
}


def config_instance(name, seed):
    """One seeded instance of BASELINE config `name` (dict as make_instance)."""
    cfg = CONFIGS[name]
    T, N, E, dyn = cfg['T'], cfg['N'], cfg['E'], cfg['dynamics']
    if name == 'metric':
        return make_instance(seed, T=T, N=N, E=E)
    if name == 'A':
        return make_instance(seed, T=T, N=N, E=E, lateral=(1.5, 6.0))
    if name == 'C':
        return make_instance(seed, T=T, N=N, E=E, kind='circle', moving=True)
    rng = np.random.default_rng(seed)
    dt, ref_speed, L = 0.1, 4.0, 3.0
    heading = rng.uniform(-np.pi, np.pi)
    start = np.array([rng.uniform(5, 55), rng.uniform(5, 55)])
    dirv = np.array([cos(heading), sin(heading)])
    nrm = np.array([-dirv[1], dirv[0]])
    state = np.array([start[0] + rng.normal(0, 0.3) * nrm[0], start[1] + rng.normal(0, 0.3) * nrm[1],
                      heading + rng.normal(0, 0.1)])
    nom_u = np.vstack([np.full(T, ref_speed), np.zeros(T)])
    nom_s = rollout(state, nom_u, dt, L, dyn)
    ref = np.zeros((3, T + 1))
    for t in range(T + 1):
        ref[0:2, t] = start + dirv * (ref_speed * dt * t)
        ref[2, t] = heading
    obstacles = []
    if name == 'B':
        # two 70 x 2 m walls at +-5 m from the line and 8 boxes 5 x 2 m beside it (corridor.yaml:23-33)
        mid = start + dirv * 30.0
        for side in (-1.0, 1.0):
            c = mid + nrm * (5.0 * side)
            obstacles.append(rect_vertices(c[0], c[1], 70.0, 2.0, heading))
        for _ in range(N - 2):
            c = start + dirv * rng.uniform(5, 55) + nrm * (rng.choice([-1.0, 1.0]) * rng.uniform(1.5, 3.5))
            obstacles.append(rect_vertices(c[0], c[1], 5.0, 2.0, rng.uniform(0, np.pi)))
    else:
        clear = 3.2                       # keep the start pose free (robot half diagonal + margin)
        while len(obstacles) < N:
            if name == 'D':
                r, a = 10.0 * np.sqrt(rng.uniform()), rng.uniform(0, 2 * np.pi)
                c = state[:2] + r * np.array([cos(a), sin(a)])
                rad = rng.uniform(0.3, 1.5)
            else:
                c = start + dirv * rng.uniform(-4, 36) + nrm * rng.uniform(-20, 20)
                rad = rng.uniform(0.5, 2.0)
            body = state[:2] + 1.5 * np.array([cos(state[2]), sin(state[2])])
            if np.linalg.norm(c - body) < rad + clear:
                continue
            obstacles.append(random_convex_polygon(rng, c[0], c[1], rad, int(rng.integers(3, E + 1))))
    obs = []
    for vert in obstacles:
        A, b = polygon_halfspaces(vert)
        obs.append(rdaobs(A, b, 'Rpositive', None, vert))
    return {'nom_s': nom_s, 'nom_u': nom_u, 'ref': ref, 'ref_speed': ref_speed, 'obstacles': obs, 'state': state}
