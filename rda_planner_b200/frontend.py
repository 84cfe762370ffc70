"""Batched front end: the steps either side of the solve, on the device.

  pre_process_batch        MPC.pre_process                         (mpc.py:251-291, 338-438)
  convert_obstacles_batch  MPC.convert_rda_obstacle + RDA_solver.assign_obstacle_parameter
                           (mpc.py:189-218, 440-549; rda_solver.py:483-526)
  BatchedMPC               MPC.control for B robots on one reference path (mpc.py:127-187),
                           closed loop without a host round trip

Thin wrappers over the C ABI (include/rda_b200.h, "front end"); no CPU fallback."""
import ctypes as C

import numpy as np
import torch

from . import _cabi
from .rda_solver import RDA_solver


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def path_tensor(ref_path, device):
    """list of (3|4, 1) waypoints (the reference's ref_path) or an array [P, >=3] -> float32 [P, 3]."""
    if isinstance(ref_path, torch.Tensor):
        return ref_path.to(device=device, dtype=torch.float32)[:, :3].contiguous()
    if isinstance(ref_path, (list, tuple)):
        arr = np.stack([np.asarray(p, float).reshape(-1)[:3] for p in ref_path])
    else:
        arr = np.asarray(ref_path, float)[:, :3]
    return torch.as_tensor(arr, dtype=torch.float32, device=device).contiguous()


def pack_shapes(obstacle_lists, max_shapes=None, max_edge_num=_cabi.MAX_EDGE):
    """Per-instance lists of simulator obstacles (attributes cone_type, center, radius, vertex,
    velocity — what MPC.convert_rda_obstacle reads, mpc.py:189-208) -> dict of host arrays in
    the layout of rda_convert_obstacles.  Polygons with more than `max_edge_num` vertices are
    refused here (the kernel would write an all-zero obstacle for them)."""
    B = len(obstacle_lists)
    M = max_shapes or max(1, max(len(l) for l in obstacle_lists))
    if M > _cabi.MAX_SHAPES:
        raise ValueError(f'at most {_cabi.MAX_SHAPES} raw obstacles per instance')
    out = {'kind': np.zeros((B, M), np.int32), 'nv': np.zeros((B, M), np.int32),
           'xy': np.zeros((B, M, _cabi.MAX_EDGE, 2), np.float32), 'radius': np.zeros((B, M), np.float32),
           'vel': np.zeros((B, M, 2), np.float32), 'count': np.zeros(B, np.int32)}
    for b, lst in enumerate(obstacle_lists):
        if len(lst) > M:
            raise ValueError('more obstacles than max_shapes')
        out['count'][b] = len(lst)
        for j, o in enumerate(lst):
            vel = getattr(o, 'velocity', None)
            if vel is not None:
                out['vel'][b, j] = np.asarray(vel, float).reshape(-1)[:2]
            if o.cone_type == 'norm2':
                out['kind'][b, j] = _cabi.OBS_CIRCLE
                out['xy'][b, j, 0] = np.asarray(o.center, float).reshape(-1)[:2]
                out['radius'][b, j] = o.radius
            else:
                v = np.asarray(o.vertex, float)
                n = v.shape[1]
                if n > min(max_edge_num, _cabi.MAX_EDGE) or n < 3:
                    raise ValueError(f'polygon with {n} vertices: 3..{min(max_edge_num, _cabi.MAX_EDGE)} supported')
                out['kind'][b, j] = _cabi.OBS_POLYGON
                out['nv'][b, j] = n
                out['xy'][b, j, :n] = v[0:2].T
    return out


def pre_process_batch(state, cur_vel, ref_speed, path, start_index, dynamics, dt, wheelbase, T,
                      threshold=0.1, ind_range=10):
    """CUDA tensors: state [B,3], cur_vel [B,2,T], ref_speed [B], path [P,3], start_index [B] int32.
    Returns nom_s [B,3,T+1], ref_s [B,3,T+1], near_index [B] int32."""
    lib = _cabi.load()
    dev = state.device
    B = state.shape[0]
    nom = torch.empty((B, 3, T + 1), dtype=torch.float32, device=dev)
    ref = torch.empty((B, 3, T + 1), dtype=torch.float32, device=dev)
    near = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(lib.rda_pre_process(B, T, _cabi.DYNAMICS[dynamics], dt, wheelbase, _ptr(state), _ptr(cur_vel),
                                        _ptr(ref_speed), _ptr(path), path.shape[0], _ptr(start_index), threshold,
                                        ind_range, _ptr(nom), _ptr(ref), _ptr(near), _stream(dev)), 'rda_pre_process')
    return nom, ref, near


def convert_obstacles_batch(shapes, state, N, T, E, dt, time_varying=False, order=True):
    """shapes: dict of CUDA tensors (kind, nv [B,M] int32; xy [B,M,8,2]; radius [B,M]; vel [B,M,2];
    count [B] int32).  Returns obs_A [B,N,Tc,E,2], obs_b [B,N,Tc,E], obs_kind [B,N], obs_count [B]."""
    lib = _cabi.load()
    dev = shapes['kind'].device
    B, M = shapes['kind'].shape
    Tc = T + 1 if time_varying else 1
    A = torch.empty((B, N, Tc, E, 2), dtype=torch.float32, device=dev)
    b = torch.empty((B, N, Tc, E), dtype=torch.float32, device=dev)
    kind = torch.empty((B, N), dtype=torch.int32, device=dev)
    count = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(lib.rda_convert_obstacles(B, M, N, T, E, dt, int(time_varying), int(order), _ptr(state),
                                              _ptr(shapes['kind']), _ptr(shapes['nv']), _ptr(shapes['xy']),
                                              _ptr(shapes['radius']), _ptr(shapes['vel']), _ptr(shapes['count']),
                                              _ptr(A), _ptr(b), _ptr(kind), _ptr(count), _stream(dev)),
                    'rda_convert_obstacles')
    return A, b, kind, count


def shapes_to_device(shapes, device):
    return {k: torch.as_tensor(v, device=device).contiguous() for k, v in shapes.items()}


class BatchedMPC:
    """MPC.control (mpc.py:127-187) for `batch` robots that follow one reference path, every step on
    the device: pre_process -> obstacle conversion -> ADMM solve -> arrive rule.  Keyword set of the
    reference's MPC where it applies.  With `enable_reverse` the path carries a gear flag in its 4th row
    (+1 forward, -1 reverse; mpc.py:139-144): it is cut into single-gear curves (split_path, mpc.py:232-249),
    every robot follows its own curve, the solver's reference speed carries the gear's sign and a robot moves on
    to the next curve when it reaches the end of the current one (mpc.py:166-183)."""

    def __init__(self, car_tuple, ref_path, batch, receding=10, sample_time=0.1, iter_num=4,
                 enable_reverse=False, obstacle_order=True, max_edge_num=5, max_obs_num=5,
                 accelerated=True, goal_index_threshold=1, device=None, iter_threshold=0.2, **kwargs):
        self.lib = _cabi.load()
        self.enable_reverse = bool(enable_reverse)
        self.rda = RDA_solver(receding, car_tuple, max_edge_num, max_obs_num, iter_num=iter_num,
                              step_time=sample_time, iter_threshold=iter_threshold, accelerated=accelerated,
                              time_print=False, batch=batch, device=device, **kwargs)
        self.device = self.rda.device
        self.batch, self.T, self.dt = batch, receding, sample_time
        self.N, self.E = max_obs_num, self.rda.max_edge_num
        self.car_tuple = car_tuple
        self.dynamics, self.L = car_tuple.dynamics, float(car_tuple.wheelbase)
        self.obstacle_order = obstacle_order
        self.goal_index_threshold = goal_index_threshold
        self.path = path_tensor(ref_path, self.device)
        self.cur_index = torch.zeros(batch, dtype=torch.int32, device=self.device)
        if self.enable_reverse:
            # split_path (mpc.py:232-249): a new curve starts wherever the gear flag changes
            flags = np.array([float(np.asarray(p, float).reshape(-1)[-1]) for p in ref_path])
            cuts = [0] + [i for i in range(1, len(flags)) if flags[i] != flags[i - 1]] + [len(flags)]
            self.curve_start = torch.as_tensor(cuts, dtype=torch.int32, device=self.device)
            self.curve_gear = torch.as_tensor([flags[c] for c in cuts[:-1]], dtype=torch.float32, device=self.device)
            self.n_curves = len(cuts) - 1
            self.curve_index = torch.zeros(batch, dtype=torch.int32, device=self.device)
        init_vel = kwargs.get('init_vel')
        self.cur_vel = torch.zeros((batch, 2, receding), dtype=torch.float32, device=self.device)
        if init_vel is not None:
            self.cur_vel[:] = torch.as_tensor(init_vel, dtype=torch.float32, device=self.device)
        self.arrive = torch.zeros(batch, dtype=torch.int32, device=self.device)
        self._empty = None

    def _no_obstacles(self):
        if self._empty is None:
            B, N, E, dev = self.batch, max(self.N, 1), self.E, self.device
            self._empty = (torch.zeros((B, N, 1, E, 2), dtype=torch.float32, device=dev),
                           torch.zeros((B, N, 1, E), dtype=torch.float32, device=dev),
                           torch.zeros((B, N), dtype=torch.int32, device=dev),
                           torch.zeros(B, dtype=torch.int32, device=dev))
        return self._empty

    def control(self, state, ref_speed=5.0, shapes=None, time_varying=False):
        """state [B,3] (CUDA tensor or array), ref_speed scalar or [B], shapes: dict from
        pack_shapes / shapes_to_device (None: free space).  Returns (u0 [B,2], info) where info holds
        the solver's batched outputs plus 'arrive', 'nom_s', 'ref_s', 'cur_index'.  No host sync."""
        dev, B, T = self.device, self.batch, self.T
        state = torch.as_tensor(state, dtype=torch.float32, device=dev).reshape(B, -1)[:, :3].contiguous()
        if not isinstance(ref_speed, torch.Tensor):
            ref_speed = torch.full((B,), float(ref_speed), dtype=torch.float32, device=dev) if np.isscalar(ref_speed) \
                else torch.as_tensor(ref_speed, dtype=torch.float32, device=dev)
        ref_speed = ref_speed.to(dtype=torch.float32).contiguous()
        solver_speed = ref_speed
        if self.enable_reverse:
            nom_s = torch.empty((B, 3, T + 1), dtype=torch.float32, device=dev)
            ref_s = torch.empty((B, 3, T + 1), dtype=torch.float32, device=dev)
            near = torch.empty(B, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                _cabi.check(self.lib.rda_pre_process_curves(
                    B, T, _cabi.DYNAMICS[self.dynamics], self.dt, self.L, _ptr(state), _ptr(self.cur_vel), _ptr(ref_speed),
                    _ptr(self.path), self.n_curves, _ptr(self.curve_start), _ptr(self.curve_index), _ptr(self.cur_index),
                    0.1, 10, _ptr(nom_s), _ptr(ref_s), _ptr(near), _stream(dev)), 'rda_pre_process_curves')
            gear = self.curve_gear[self.curve_index.long().clamp(max=self.n_curves - 1)]
            solver_speed = (ref_speed * gear).contiguous()                      # gear_flag * ref_speed (mpc.py:161)
        else:
            nom_s, ref_s, near = pre_process_batch(state, self.cur_vel, ref_speed, self.path, self.cur_index,
                                                   self.dynamics, self.dt, self.L, T)
        self.cur_index = near
        if shapes is None or self.N == 0:
            A, b, kind, count = self._no_obstacles()
            time_varying = False
        else:
            A, b, kind, count = convert_obstacles_batch(shapes, state, self.N, T, self.E, self.dt, time_varying,
                                                        self.obstacle_order)
        out = self.rda.iterative_solve_batch(nom_s, self.cur_vel, ref_s, solver_speed, A, b, kind, count, time_varying)
        with torch.cuda.device(dev):
            if self.enable_reverse:
                _cabi.check(self.lib.rda_post_process_gear(B, T, self.n_curves, _ptr(self.curve_start), self.goal_index_threshold,
                                                           _ptr(near), _ptr(self.curve_index), _ptr(out['u']), _ptr(self.cur_vel),
                                                           _ptr(self.arrive), _stream(dev)), 'rda_post_process_gear')
            else:
                _cabi.check(self.lib.rda_post_process(B, T, self.path.shape[0], self.goal_index_threshold, _ptr(near),
                                                      _ptr(out['u']), _ptr(self.cur_vel), _ptr(self.arrive), _stream(dev)),
                            'rda_post_process')
        info = dict(out)
        info.update(arrive=self.arrive, nom_s=nom_s, ref_s=ref_s, cur_index=near)
        if self.enable_reverse:
            info['curve_index'] = self.curve_index
        return out['u'][:, :, 0], info

    def advance(self, state):
        """One simulator step with the first control of the last solve (mpc.py:293-336), in place."""
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.rda_motion_predict(self.batch, self.T, _cabi.DYNAMICS[self.dynamics], self.dt, self.L,
                                                    _ptr(self.cur_vel), _ptr(state), _stream(self.device)),
                        'rda_motion_predict')
        return state

    def reset(self):
        self.rda.reset()
        self.cur_index.zero_()
        self.cur_vel.zero_()
        if self.enable_reverse:
            self.curve_index.zero_()
