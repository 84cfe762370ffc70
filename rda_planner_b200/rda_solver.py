"""RDA_solver — the drop-in for RDA_planner.rda_solver.RDA_solver, backed by sm_100a kernels.

Mirror of /root/reference/RDA_planner/rda_solver.py (class RDA_solver): constructor :18-22
(+ tunables :185-201, ws/wu :218-219), iterative_solve :573-610, assign_adjust_parameter
:426-434, get_adjust_parameter :1055-1056, reset :1060-1068.  numpy in / numpy out at this
level, exactly like the reference; all arithmetic runs in librda_b200.so (include/rda_b200.h)
on the current CUDA device, with PyTorch used only for device memory and streams.

New surface (absent in the reference, required by BASELINE.json): `batch` > 1 and
`iterative_solve_batch` on CUDA tensors — B independent planning instances that share
(T, N, E, robot, dynamics) and keep their own warm-start state on the device.
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _cabi


def _as_cuda_f32(x, device):
    return torch.as_tensor(x, dtype=torch.float32, device=device).contiguous()


# Half-size [m] of the square (centred on the robot) that closes unbounded half-space sets in canonical_polygon_rows.
# The kernels work in float32 relative to the robot: a vertex L metres away turns a direction error of one ulp into
# L * 1e-7 m of margin, so the square is kept as small as a planning horizon allows and grown only when the set does
# not reach into it (then the obstacle is too far to matter).
HALFSPACE_BOUND = 100.0
HALFSPACE_BOUND_MAX = 1.0e5
_CANONICAL_SEEN = set()      # (A bytes, b bytes) of row sets already verified to be closed counter-clockwise polygons


def canonical_polygon_rows(A, b, bound=None, center=None):
    """Rows of a convex set {x: Ax <= b} ordered counter-clockwise by normal angle so that vertex i joins rows i-1 and
    i — the layout the kernels' closest-point geometry needs.  A no-op for the output of mpc.py:476-510 (closed convex
    polygon, rows already in order).  Anything else the reference accepts through rda_obstacle=True (mpc.py:150-155:
    the caller's own (A, b) tuples) is reduced to that case by clipping: rows in any order, redundant rows (dropped:
    their multipliers are zero at every optimum), and UNBOUNDED sets (a wall, a wedge, a strip), which are closed by
    the sides of the square |x - center|_inf <= bound (center: the robot position, default the origin; bound: default
    HALFSPACE_BOUND, grown 4x until the set reaches into the square) — exact as long as the closest obstacle point to
    the robot does not lie on one of those artificial sides.  Original rows keep their scaling.  Raises ValueError for
    an empty set."""
    A = np.asarray(A, float)
    b = np.asarray(b, float).reshape(-1)
    n = A.shape[0]
    def one_turn(M):
        # consecutive normals turn left AND the turning angles add up to a single 2 pi (no double winding)
        m = M.shape[0]
        d = M[np.arange(m) - 1, 0] * M[:, 1] - M[np.arange(m) - 1, 1] * M[:, 0]
        if m < 3 or not np.all(d > 0):
            return False
        a = np.arctan2(M[:, 1], M[:, 0])
        turn = np.mod(a - np.roll(a, 1), 2 * np.pi)
        return abs(turn.sum() - 2 * np.pi) < 1e-6
    def every_row_is_an_edge(M, c):
        # vertex i = rows i-1 and i; every vertex must satisfy all rows, and consecutive vertices must differ
        Mp, cp = np.roll(M, 1, axis=0), np.roll(c, 1)
        det = Mp[:, 0] * M[:, 1] - Mp[:, 1] * M[:, 0]
        V = np.stack([(cp * M[:, 1] - c * Mp[:, 1]) / det, (Mp[:, 0] * c - M[:, 0] * cp) / det], axis=1)
        slack = c[None, :] - V @ M.T
        scale = np.linalg.norm(M, axis=1)[None, :] * (1.0 + np.abs(V).max())
        if np.any(slack < -1e-9 * scale):
            return False
        return bool(np.all(np.linalg.norm(np.roll(V, -1, axis=0) - V, axis=1) > 1e-9 * (1.0 + np.abs(V).max())))
    # the common case — the output of mpc.py:476-510, unchanged from one control step to the next for static obstacles — is
    # recognised once and remembered by content (a dozen small numpy calls per obstacle would otherwise cost more than the
    # whole solve of the path_track example)
    key = (A.tobytes(), b.tobytes()) if A.size <= 64 else None
    if key is not None and key in _CANONICAL_SEEN:
        return A, b
    live = np.linalg.norm(A, axis=1) > 0
    if live.all() and one_turn(A) and every_row_is_an_edge(A, b):
        if key is not None:
            if len(_CANONICAL_SEEN) > 65536:
                _CANONICAL_SEEN.clear()
            _CANONICAL_SEEN.add(key)
        return A, b
    if np.any(b[~live] < 0):
        raise ValueError('obstacle half-spaces describe an empty set (0 <= b violated by a zero row)')
    ctr = np.zeros(2) if center is None else np.asarray(center, float).reshape(-1)[:2]
    if bound is None or bound > 0:
        # first square of half-size `bound` (default HALFSPACE_BOUND), grown 4x while the set does not reach into it
        L = float(HALFSPACE_BOUND if bound is None else bound)
        while True:
            try:
                return canonical_polygon_rows(A, b, bound=-L, center=ctr)
            except ValueError:
                if L >= HALFSPACE_BOUND_MAX:
                    raise
                L *= 4.0
    L = -float(bound)            # negative: exactly this size, no growth (internal)
    # Sutherland-Hodgman clipping of the square by every row; each vertex carries the row of the edge that STARTS there
    # (-1..-4: the artificial sides of the square)
    poly = [(ctr + np.array([-L, -L]), -1), (ctr + np.array([L, -L]), -2), (ctr + np.array([L, L]), -3),
            (ctr + np.array([-L, L]), -4)]
    for r in np.nonzero(live)[0]:
        a, c = A[r], b[r]
        tol = 1e-12 * np.linalg.norm(a) * (L + np.abs(ctr).max())
        out = []
        for k in range(len(poly)):
            (P, tag), (Q, _) = poly[k], poly[(k + 1) % len(poly)]
            sp, sq = a @ P - c, a @ Q - c
            if sp <= tol:
                out.append((P, tag))
                if sq > tol:
                    out.append((P + (Q - P) * (sp / (sp - sq)), int(r)))
            elif sq <= tol:
                out.append((P + (Q - P) * (sp / (sp - sq)), tag))
        poly = out
        if len(poly) < 3:
            raise ValueError('obstacle half-spaces describe an empty set inside |x| <= %g' % L)
    # drop zero-length edges (a row that only touches a vertex)
    keep = [k for k in range(len(poly))
            if np.linalg.norm(poly[(k + 1) % len(poly)][0] - poly[k][0]) > 1e-9 * (1.0 + 1e-4 * (L + np.abs(ctr).max()))]
    poly = [poly[k] for k in keep]
    if len(poly) < 3:
        raise ValueError('obstacle half-spaces describe a set without interior')
    box = {-1: (np.array([0.0, -1.0]), L - ctr[1]), -2: (np.array([1.0, 0.0]), L + ctr[0]),
           -3: (np.array([0.0, 1.0]), L + ctr[1]), -4: (np.array([-1.0, 0.0]), L - ctr[0])}
    rows = [(A[t], b[t]) if t >= 0 else box[t] for _, t in poly]
    An = np.array([r[0] for r in rows], float)
    bn = np.array([r[1] for r in rows], float)
    if not one_turn(An):
        raise ValueError('obstacle half-spaces could not be reduced to a closed convex polygon')
    return An, bn


def pack_obstacles(obstacle_list, T, N, E, center=None, bound=None):
    """assign_obstacle_parameter (rda_solver.py:483-526): pad a short list by repeating its
    last element (mutating the caller's list, as the reference does), truncate a long one,
    zero-pad rows to E.  center: robot position, only used to close unbounded half-space sets
    (canonical_polygon_rows).  Returns (A [N,Tc,E,2], b [N,Tc,E], kind [N], count, time_varying)."""
    count = len(obstacle_list)
    if 0 < count < N:
        obstacle_list += [obstacle_list[-1]] * (N - count)
    number = min(len(obstacle_list), N)
    tv = any(isinstance(o.A, list) for o in obstacle_list[:number])
    Tc = T + 1 if tv else 1
    A = np.zeros((N, Tc, E, 2), np.float32)
    b = np.zeros((N, Tc, E), np.float32)
    kind = np.zeros(N, np.int32)
    for i in range(number):
        o = obstacle_list[i]
        circle = o.cone_type != 'Rpositive'
        kind[i] = _cabi.OBS_CIRCLE if circle else _cabi.OBS_POLYGON
        for t in range(Tc):
            At = o.A[t] if isinstance(o.A, list) else o.A
            bt = o.b[t] if isinstance(o.b, list) else o.b
            At = np.asarray(At, float)
            bt = np.asarray(bt, float).reshape(-1)
            if not circle:
                At, bt = canonical_polygon_rows(At, bt, bound=bound, center=center)
            en = At.shape[0]
            if en > E:
                raise ValueError(f'obstacle with {en} edges exceeds max_edge_num={E}'
                                 + (' (an unbounded half-space set is closed by up to four sides of a large square: '
                                    'raise max_edge_num accordingly)' if en > np.asarray(o.A[t] if isinstance(o.A, list) else o.A).shape[0] else ''))
            A[i, t, :en] = At
            b[i, t, :en] = bt
    return A, b, kind, count, tv


class RDA_solver:
    def __init__(self, receding, car_tuple, max_edge_num=5, max_obs_num=5, iter_num=2, step_time=0.1,
                 iter_threshold=0.2, process_num=4, accelerated=True, time_print=True, batch=1,
                 device=None, su_fp64=True, z_theta=0.5, graph=False, **kwargs):
        """kwargs: slack_gain (8), max_sd (1.0), min_sd (0.1), ro1 (200), ro2 (1), ws (1), wu (1)
        — rda_solver.py:24-31.  `process_num` is accepted for compatibility and ignored.
        graph=True replays the launches of a solve from a CUDA graph (captured on first use per set of input
        buffers / iteration count): the single-instance API then copies its inputs into persistent device
        buffers, so every control step after the first is one graph launch."""
        if not torch.cuda.is_available():
            raise RuntimeError('rda_planner_b200 needs a CUDA device (B200); there is no CPU fallback')
        self.lib = _cabi.load()
        self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        self.T = receding
        self.car_tuple = car_tuple
        self.L = car_tuple.wheelbase
        self.max_obs_num = max_obs_num
        self.max_edge_num = max(max_edge_num, 3)
        self.dynamics = car_tuple.dynamics
        self.iter_num = iter_num
        self.dt = step_time
        self.iter_threshold = iter_threshold
        self.accelerated = accelerated
        self.process_num = process_num
        self.time_print = time_print
        self.batch = batch
        self.ws = kwargs.get('ws', 1)
        self.wu = kwargs.get('wu', 1)
        G = np.asarray(car_tuple.G, float)
        h = np.asarray(car_tuple.h, float).reshape(-1)
        if car_tuple.cone_type == 'norm2':
            # disc body (cone_cp_array(-mu, 'norm2'), :1034-1039) as ir-sim describes it: |y - (h0, h1)| <= -h2
            if G.shape != (3, 2) or np.abs(G - np.array([[1.0, 0], [0, 1], [0, 0]])).max() > 1e-9 or not h[2] < 0:
                raise NotImplementedError("norm2 robot: only the disc G = [[1,0],[0,1],[0,0]], h = (cx, cy, -r) is supported")
            robot_cone = _cabi.ROBOT_DISC
        elif car_tuple.cone_type == 'Rpositive':
            G, h = canonical_polygon_rows(G, h)
            robot_cone = _cabi.ROBOT_POLYGON
        else:
            raise ValueError(f'unknown robot cone type {car_tuple.cone_type!r}')
        R = G.shape[0]
        if R > _cabi.MAX_ROBOT_EDGE or self.max_edge_num > _cabi.MAX_EDGE:
            raise ValueError('at most 8 robot edges / obstacle edges are supported')
        cfg = _cabi.Config()
        cfg.batch, cfg.receding, cfg.max_obs_num = batch, receding, max_obs_num
        cfg.max_edge_num, cfg.robot_edges = self.max_edge_num, R
        cfg.dynamics = _cabi.DYNAMICS[self.dynamics]
        cfg.accelerated = int(bool(accelerated))
        cfg.su_fp64 = int(bool(su_fp64))
        cfg.step_time, cfg.wheelbase = step_time, float(self.L)
        ms = np.asarray(car_tuple.max_speed, float).reshape(-1)
        ma = np.asarray(car_tuple.max_acce, float).reshape(-1)
        for k in range(2):
            cfg.max_speed[k] = ms[k]
            cfg.acce_bound[k] = ma[k] * step_time                      # :44
        cfg.ws, cfg.wu = self.ws, self.wu
        cfg.robot_cone = robot_cone
        for j in range(R):
            cfg.G[2 * j], cfg.G[2 * j + 1], cfg.h[j] = G[j, 0], G[j, 1], h[j]
        self._tun = _cabi.Tunables(kwargs.get('slack_gain', 8), kwargs.get('max_sd', 1.0),
                                   kwargs.get('min_sd', 0.1), kwargs.get('ro1', 200),
                                   kwargs.get('ro2', 1), z_theta)
        # the values as the caller gave them (the device copy is float32): get_adjust_parameter returns these
        self._tun_py = {k: kwargs.get(k, dflt) for k, dflt in (('slack_gain', 8), ('max_sd', 1.0), ('min_sd', 0.1),
                                                                ('ro1', 200), ('ro2', 1))}
        # closing square of unbounded half-space obstacles (canonical_polygon_rows): 1.5 x what the robot can reach within the
        # horizon plus body and safety distance, at least 30 m — small, because far vertices cost float32 precision
        self._hs_bound = max(30.0, 1.5 * (receding * step_time * float(abs(ms[0])) + 8.0))
        self._cfg = cfg
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.rda_create(C.byref(cfg), C.byref(self._tun), C.byref(self._h)), 'rda_create')
        B, T = batch, receding
        dev = self.device
        self._out = {
            'u': torch.empty((B, 2, T), dtype=torch.float32, device=dev),
            's': torch.empty((B, 3, T + 1), dtype=torch.float32, device=dev),
            'resi_pri': torch.empty(B, dtype=torch.float32, device=dev),
            'resi_dual': torch.empty(B, dtype=torch.float32, device=dev),
            'status': torch.empty(B, dtype=torch.int32, device=dev),
            'iters': torch.empty(B, dtype=torch.int32, device=dev),
        }
        self._keep = None
        self.obstacle_num = 0
        self.use_graph = bool(graph)
        self._graphs = {}
        self._static = None

    def __del__(self):
        try:
            if getattr(self, '_h', None) is not None and self._h.value:
                self.lib.rda_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ tunables
    def assign_adjust_parameter(self, **kwargs):
        """slack_gain, max_sd, min_sd, ro1, ro2 (ws/wu silently ignored, as in :426-434)."""
        t = self._tun
        t.slack_gain = kwargs.get('slack_gain', t.slack_gain)
        t.max_sd = kwargs.get('max_sd', t.max_sd)
        t.min_sd = kwargs.get('min_sd', t.min_sd)
        t.ro1 = kwargs.get('ro1', t.ro1)
        t.ro2 = kwargs.get('ro2', t.ro2)
        for k in self._tun_py:
            if k in kwargs:
                self._tun_py[k] = kwargs[k]
        _cabi.check(self.lib.rda_set_tunables(self._h, C.byref(t)), 'rda_set_tunables')

    def get_adjust_parameter(self):
        t = _cabi.Tunables()
        _cabi.check(self.lib.rda_get_tunables(self._h, C.byref(t)), 'rda_get_tunables')
        out = dict(self._tun_py)          # exact Python values, as the reference returns them (:1055-1056)
        for k in out:                      # ... unless the handle was changed behind this object's back
            if abs(float(getattr(t, k)) - float(out[k])) > 1e-6 * (1 + abs(float(out[k]))):
                out[k] = float(getattr(t, k))
        out.update(ws=self.ws, wu=self.wu)
        return out

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self):
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.rda_reset(self._h, self._stream()), 'rda_reset')

    def cold_start(self):
        """Extension: forget every warm-start quantity (constructor state)."""
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.rda_cold_start(self._h, self._stream()), 'rda_cold_start')

    def buffer_count(self, buf_id):
        ptr, cnt = C.c_void_p(), C.c_size_t()
        _cabi.check(self.lib.rda_get_buffer(self._h, buf_id, C.byref(ptr), C.byref(cnt)), 'rda_get_buffer')
        return cnt.value

    def state_buffer(self, buf_id, shape=None):
        """Copy of a persistent device buffer (checkpointing / tests)."""
        dtype = torch.int32 if buf_id == _cabi.BUF_COUNTERS else torch.float32
        out = torch.empty(self.buffer_count(buf_id), dtype=dtype, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.rda_copy_buffer(self._h, buf_id, out.data_ptr(), 0, self._stream()),
                        'rda_copy_buffer')
        return out if shape is None else out.reshape(shape)

    def load_state_buffer(self, buf_id, values):
        """Overwrite a persistent device buffer (resume / tests)."""
        dtype = torch.int32 if buf_id == _cabi.BUF_COUNTERS else torch.float32
        src = torch.as_tensor(values, dtype=dtype, device=self.device).contiguous().reshape(-1)
        if src.numel() != self.buffer_count(buf_id):
            raise ValueError('wrong element count')
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.rda_copy_buffer(self._h, buf_id, src.data_ptr(), 1, self._stream()),
                        'rda_copy_buffer')
        self._keep_state = src

    # ------------------------------------------------------------------ solve
    def _inputs(self, nom_s, nom_u, ref_s, ref_speed, obs_A, obs_b, obs_kind, obs_count, time_varying):
        B, T, N, E = self.batch, self.T, self.max_obs_num, self.max_edge_num
        dev = self.device
        t = {
            'nom_s': _as_cuda_f32(nom_s, dev).reshape(B, 3, T + 1),
            'nom_u': _as_cuda_f32(nom_u, dev).reshape(B, 2, T),
            'ref_s': _as_cuda_f32(ref_s, dev).reshape(B, 3, T + 1),
            'ref_speed': _as_cuda_f32(ref_speed, dev).reshape(B),
        }
        Tc = T + 1 if time_varying else 1
        if N > 0:
            t['obs_A'] = _as_cuda_f32(obs_A, dev).reshape(B, N, Tc, E, 2)
            t['obs_b'] = _as_cuda_f32(obs_b, dev).reshape(B, N, Tc, E)
            t['obs_kind'] = torch.as_tensor(obs_kind, dtype=torch.int32, device=dev).reshape(B, N).contiguous()
            t['obs_count'] = torch.as_tensor(obs_count, dtype=torch.int32, device=dev).reshape(B).contiguous()
        inp = _cabi.Inputs()
        for k in ('nom_s', 'nom_u', 'ref_s', 'ref_speed', 'obs_A', 'obs_b', 'obs_kind', 'obs_count'):
            setattr(inp, k, t[k].data_ptr() if k in t else None)
        inp.obs_time_varying = int(bool(time_varying))
        self._keep = t          # the kernels read these buffers asynchronously
        return inp

    def _outputs(self):
        o = _cabi.Outputs()
        o.u_opt, o.s_opt = self._out['u'].data_ptr(), self._out['s'].data_ptr()
        o.resi_pri, o.resi_dual = self._out['resi_pri'].data_ptr(), self._out['resi_dual'].data_ptr()
        o.status, o.iters = self._out['status'].data_ptr(), self._out['iters'].data_ptr()
        return o

    def iterative_solve_batch(self, nom_s, nom_u, ref_s, ref_speed, obs_A=None, obs_b=None, obs_kind=None,
                              obs_count=None, time_varying=False, iter_num=None, iter_threshold=None):
        """B instances at once.  Arguments are array-likes or CUDA tensors with the layouts of
        include/rda_b200.h; returns a dict of CUDA tensors (u [B,2,T], s [B,3,T+1], resi_pri,
        resi_dual, status, iters) valid on the current stream.  No host synchronisation."""
        inp = self._inputs(nom_s, nom_u, ref_s, ref_speed, obs_A, obs_b, obs_kind, obs_count, time_varying)
        out = self._outputs()
        it = self.iter_num if iter_num is None else iter_num
        thr = self.iter_threshold if iter_threshold is None else iter_threshold
        if self.use_graph:
            # one graph per (input buffers, iteration count, threshold): the launches read the inputs through
            # their addresses, so a replay is valid exactly while the caller reuses the same tensors
            key = (tuple(int(getattr(inp, k) or 0) for k in ('nom_s', 'nom_u', 'ref_s', 'ref_speed', 'obs_A', 'obs_b',
                                                              'obs_kind', 'obs_count')), int(inp.obs_time_varying), int(it), float(thr))
            g = self._graphs.get(key)
            if g is None:
                cur = torch.cuda.current_stream(self.device)
                side = torch.cuda.Stream(self.device)
                side.wait_stream(cur)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g, stream=side):
                        _cabi.check(self.lib.rda_solve(self._h, C.byref(inp), C.byref(out), int(it), float(thr),
                                                       self._stream()), 'rda_solve (graph capture)')
                cur.wait_stream(side)
                if len(self._graphs) >= 8:
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[key] = (g, self._keep)          # keep the captured input tensors alive
                g = self._graphs[key]
            g[0].replay()
            return self._out
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.rda_solve(self._h, C.byref(inp), C.byref(out), int(it), float(thr),
                                           self._stream()), 'rda_solve')
        return self._out

    # phase API (rda_begin / rda_step_su / rda_step_lammuz / rda_finish): unit tests, profiling
    def begin(self, nom_s, nom_u, ref_s, ref_speed, obs_A=None, obs_b=None, obs_kind=None, obs_count=None,
              time_varying=False, iter_threshold=None):
        inp = self._inputs(nom_s, nom_u, ref_s, ref_speed, obs_A, obs_b, obs_kind, obs_count, time_varying)
        thr = self.iter_threshold if iter_threshold is None else iter_threshold
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.rda_begin(self._h, C.byref(inp), float(thr), self._stream()), 'rda_begin')

    def step_su(self):
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.rda_step_su(self._h, self._stream()), 'rda_step_su')

    def step_lammuz(self):
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.rda_step_lammuz(self._h, self._stream()), 'rda_step_lammuz')

    def finish(self):
        out = self._outputs()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.rda_finish(self._h, C.byref(out), self._stream()), 'rda_finish')
        return self._out

    def launch_count(self):
        return self.lib.rda_last_launch_count(self._h)

    def _solve_static(self, nom_s, nom_u, ref, ref_speed, A, b, kind, count, tv):
        """graph=True, single instance: inputs are staged into persistent device buffers (one pinned host block, one
        H2D copy each) so that the captured graph of the solve can be replayed every control step."""
        T, N, E, dev = self.T, self.max_obs_num, self.max_edge_num, self.device
        Tc = T + 1 if tv else 1
        if self._static is None or self._static['Tc'] != Tc:
            z = lambda *sh, dt=torch.float32: torch.zeros(sh, dtype=dt, device=dev)
            self._static = {'Tc': Tc, 'nom_s': z(1, 3, T + 1), 'nom_u': z(1, 2, T), 'ref_s': z(1, 3, T + 1), 'ref_speed': z(1),
                            'obs_A': z(1, max(N, 1), Tc, E, 2), 'obs_b': z(1, max(N, 1), Tc, E),
                            'obs_kind': z(1, max(N, 1), dt=torch.int32), 'obs_count': z(1, dt=torch.int32)}
            self._graphs.clear()
        st = self._static
        st['nom_s'].copy_(torch.as_tensor(nom_s, dtype=torch.float32)[None])
        st['nom_u'].copy_(torch.as_tensor(nom_u, dtype=torch.float32)[None])
        st['ref_s'].copy_(torch.as_tensor(ref, dtype=torch.float32)[None])
        st['ref_speed'].fill_(ref_speed)
        if N > 0:
            st['obs_A'].copy_(torch.as_tensor(A, dtype=torch.float32)[None])
            st['obs_b'].copy_(torch.as_tensor(b, dtype=torch.float32)[None])
            st['obs_kind'].copy_(torch.as_tensor(kind, dtype=torch.int32)[None])
        st['obs_count'].fill_(int(count))
        return self.iterative_solve_batch(st['nom_s'], st['nom_u'], st['ref_s'], st['ref_speed'],
                                          st['obs_A'] if N > 0 else None, st['obs_b'] if N > 0 else None,
                                          st['obs_kind'] if N > 0 else None, st['obs_count'] if N > 0 else None, tv)

    def iterative_solve(self, nom_s, nom_u, ref_states, ref_speed, obstacle_list, **kwargs):
        """Reference signature (:573): numpy in, (u (2,T) ndarray, info dict) out."""
        if self.batch != 1:
            raise RuntimeError('iterative_solve is the single-instance API; use iterative_solve_batch')
        T, N, E = self.T, self.max_obs_num, self.max_edge_num
        start = time.time()
        ref = np.hstack(ref_states)[0:3, :]                                     # :580
        if N > 0:
            A, b, kind, count, tv = pack_obstacles(obstacle_list, T, N, E, center=np.asarray(nom_s, float)[0:2, 0],
                                                   bound=self._hs_bound)
        else:
            A = b = kind = None
            count, tv = len(obstacle_list), False
        self.obstacle_num = len(obstacle_list)
        if self.use_graph:
            res = self._solve_static(np.asarray(nom_s, float), np.asarray(nom_u, float), ref, float(ref_speed), A, b, kind,
                                     count, tv)
        else:
            res = self.iterative_solve_batch(np.asarray(nom_s, float)[None], np.asarray(nom_u, float)[None],
                                             ref[None], np.array([ref_speed], float),
                                             None if A is None else A[None], None if b is None else b[None],
                                             None if kind is None else kind[None], np.array([count]), tv)
        u = res['u'][0].double().cpu().numpy()
        s = res['s'][0].double().cpu().numpy()
        info = {'ref_traj_list': ref_states, 'opt_state_list': [s[:, t:t + 1] for t in range(T + 1)],
                'iteration_time': time.time() - start,
                'resi_dual': float(res['resi_dual'][0]), 'resi_pri': float(res['resi_pri'][0]),
                'status': int(res['status'][0]), 'iterations': int(res['iters'][0])}
        if self.time_print:
            print('iterations:', info['iterations'], ' time:', info['iteration_time'])
        return u, info
