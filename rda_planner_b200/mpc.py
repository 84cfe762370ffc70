"""Host-side MPC front end (Python, numpy) — the drop-in for RDA_planner.mpc.MPC.

Behavioural mirror of /root/reference/RDA_planner/mpc.py (class MPC :15-569): same
constructor keywords (:67-84), same `control()` contract (:127-187), same obstacle
conversion (:189-218, :440-549) and reference pre-processing (:251-423).  The only
thing it does with the result of the pre-processing is to call
`self.rda.iterative_solve(...)` (:157-164), which here is the B200 CUDA solver of
rda_planner_b200.rda_solver.RDA_solver.  This module is scalar host logic, outside
the kernel hot path.
"""
from collections import namedtuple
from math import sqrt, pi, sin, cos, tan, inf

import numpy as np

rdaobs = namedtuple("rdaobs", "A b cone_type center vertex")   # mpc.py:12


def wrap_to_pi(angle):
    """Bring an angle into [-pi, pi] by whole turns (mpc.py:431-438)."""
    while angle > pi:
        angle -= 2 * pi
    while angle < -pi:
        angle += 2 * pi
    return angle


def _turn(o, a, b):
    return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])


def polygon_order(points):
    """(is_convex, 'CW'|'CCW'|None) for a 2 x n vertex array (mpc.py:518-549)."""
    n = points.shape[1]
    if n < 3:
        return False, None
    sign = 0
    for i in range(n):
        cr = _turn(points[:, i], points[:, (i + 1) % n], points[:, (i + 2) % n])
        if cr == 0:
            continue
        if sign == 0:
            sign = 1 if cr > 0 else -1
        elif (cr > 0) != (sign > 0):
            return False, None
    return True, ("CCW" if sign > 0 else "CW")


def polygon_halfspaces(vertex):
    """Edge half-spaces A x <= b of a convex polygon given as a 2 x n vertex array;
    row i is the (un-normalised) outward normal of edge i -> i+1 after the vertices
    have been put in counter-clockwise order (mpc.py:476-510)."""
    ok, order = polygon_order(vertex)
    if not ok:
        print(f"Warning: The polygon constructed by vertex is not convex. Please check the vertex: {vertex}")
    if order == "CW":
        vertex = vertex[:, ::-1]
    nxt = np.roll(vertex[0:2], -1, axis=1)
    edge = nxt - vertex[0:2]
    A = np.stack([edge[1], -edge[0]], axis=1).astype(float)
    b = np.sum(A * vertex[0:2].T, axis=1, keepdims=True)
    return A, b


def seg_circle_exit(center, radius, seg):
    """Far intersection of the circle (center, radius) with segment seg=[p0, p1], or None
    (mpc.py:385-423)."""
    p0, p1 = seg
    assert center.shape == (2,) and p0.shape == (2,) and p1.shape == (2,)
    d = p1 - p0
    if np.linalg.norm(d) == 0:
        return None
    f = p0 - center
    qa = d @ d
    qb = 2 * f @ d
    qc = f @ f - radius ** 2
    disc = qb ** 2 - 4 * qa * qc
    if disc < 0:
        return None
    t_far = (-qb + sqrt(disc)) / (2 * qa)
    if 0 <= t_far <= 1:
        return p0 + t_far * d
    return None


class MPC:
    """See the reference docstring (mpc.py:16-65) for the meaning of every argument; the
    keyword set is identical.  `process_num` is accepted and ignored (obstacle-level
    process parallelism is replaced by thread-level parallelism on the GPU).
    `solver_cls` (extension) lets tests inject another object with the RDA_solver
    interface."""

    def __init__(self, car_tuple, ref_path, receding=10, sample_time=0.1, iter_num=4,
                 enable_reverse=False, rda_obstacle=False, obstacle_order=True,
                 max_edge_num=5, max_obs_num=5, process_num=4, accelerated=True,
                 time_print=False, goal_index_threshold=1, solver_cls=None, **kwargs):
        self.car_tuple = car_tuple
        self.L = car_tuple.wheelbase
        self.dynamics = car_tuple.dynamics
        self.receding = receding
        self.dt = sample_time
        self.cur_vel_array = kwargs.get("init_vel", np.zeros((2, receding)))
        self.state = np.zeros((3, 1))
        self.cur_index = 0
        self.ref_path = ref_path
        if solver_cls is None:
            from .rda_solver import RDA_solver as solver_cls
        self.rda = solver_cls(receding, car_tuple, max_edge_num, max_obs_num, iter_num=iter_num,
                              step_time=sample_time, process_num=process_num,
                              accelerated=accelerated, time_print=time_print, **kwargs)
        self.enable_reverse = enable_reverse
        self.rda_obstacle = rda_obstacle
        self.obstacle_order = obstacle_order
        self.goal_index_threshold = goal_index_threshold
        if enable_reverse:
            self.curve_list = self.split_path(self.ref_path)
            self.curve_index = 0

    # ------------------------------------------------------------------ control step
    def control(self, state, ref_speed=5, obstacle_list=[], **kwargs):
        """One receding-horizon step: returns (u[:, 0:1], info) (mpc.py:127-187)."""
        if np.shape(state)[0] > 3:
            state = state[0:3]
        self.state = state
        if self.enable_reverse:
            path = self.curve_list[self.curve_index]
            gear = path[0][-1, 0]
        else:
            path = self.ref_path
            gear = 1
        nom_s, ref_traj, self.cur_index = self.pre_process(state, path, self.cur_index, ref_speed, **kwargs)
        if self.rda_obstacle:
            obs = obstacle_list
        else:
            obs = self.convert_rda_obstacle(obstacle_list, state, self.obstacle_order)
        u_opt, info = self.rda.iterative_solve(nom_s, self.cur_vel_array, ref_traj, gear * ref_speed,
                                               obs, **kwargs)
        info["arrive"] = False
        if self.cur_index >= len(path) - self.goal_index_threshold:
            last_segment = True
            if self.enable_reverse:
                self.curve_index += 1
                self.cur_index = 0
                last_segment = self.curve_index >= len(self.curve_list)
            if last_segment:
                u_opt = np.zeros((2, self.receding))
                info["arrive"] = True
        self.cur_vel_array = u_opt
        return u_opt[:, 0:1], info

    # ------------------------------------------------------------------ obstacles
    def convert_rda_obstacle(self, obstacle_list, state=None, obstacle_order=False):
        """Simulator obstacles -> rdaobs(A, b, cone, center, vertex) (mpc.py:189-208)."""
        out = []
        for o in obstacle_list:
            if o.cone_type == "norm2":
                A, b = self.convert_inequal_circle(o.center, o.radius, o.velocity)
                out.append(rdaobs(A, b, o.cone_type, o.center, None))
            elif o.cone_type == "Rpositive":
                A, b = self.convert_inequal_polygon(o.vertex, o.velocity)
                out.append(rdaobs(A, b, o.cone_type, None, o.vertex))
        if obstacle_order:
            out.sort(key=self.rda_obs_distance)
        return out

    def rda_obs_distance(self, rda_obs):
        """Sort key: centre distance (disc) or nearest-vertex distance (polygon), mpc.py:210-218."""
        if rda_obs.cone_type == "norm2":
            return MPC.distance(self.state[0:2], rda_obs.center[0:2])
        if rda_obs.cone_type == "Rpositive":
            return np.min(np.linalg.norm(self.state[0:2] - rda_obs.vertex, axis=0))

    def convert_inequal_circle(self, center, radius, velocity=np.zeros((2, 1))):
        """Disc as A=[I;0], b=[c;-r] (norm2 cone); a list of T+1 copies when it moves
        faster than 0.01 (mpc.py:440-458)."""
        A0 = np.array([[1, 0], [0, 1], [0, 0]])
        tail = -radius * np.ones((1, 1))
        if np.linalg.norm(velocity) <= 0.01:
            return A0, np.vstack((center, tail))
        A, b = [], []
        for t in range(self.receding + 1):
            A.append(A0.copy())
            b.append(np.vstack((center + velocity * (t * self.dt), tail)))
        return A, b

    def convert_inequal_polygon(self, vertex, velocity=np.zeros((2, 1))):
        """Polygon half-spaces, constant-velocity copies when moving (mpc.py:460-474)."""
        if np.linalg.norm(velocity) <= 0.01:
            return self.gen_inequal_global(vertex)
        A, b = [], []
        for t in range(self.receding + 1):
            At, bt = self.gen_inequal_global(vertex + velocity * (t * self.dt))
            A.append(At)
            b.append(bt)
        return A, b

    def gen_inequal_global(self, vertex):
        return polygon_halfspaces(vertex)

    def cross_product(self, o, a, b):
        return _turn(o, a, b)

    def is_convex_and_ordered(self, points):
        return polygon_order(points)

    # ------------------------------------------------------------------ reference path
    def update_ref_path(self, ref_path):
        self.ref_path = ref_path
        self.cur_index = 0
        if self.enable_reverse:
            self.curve_list = self.split_path(self.ref_path)
            self.curve_index = 0

    def update_parameter(self, **kwargs):
        self.rda.assign_adjust_parameter(**kwargs)

    def split_path(self, ref_path):
        """Cut the path where the gear flag (last row) changes (mpc.py:232-249)."""
        pieces, start, flag = [], 0, ref_path[0][-1, 0]
        for i, pt in enumerate(ref_path):
            if pt[-1, 0] != flag:
                pieces.append(ref_path[start:i])
                start, flag = i, pt[-1, 0]
        pieces.append(ref_path[start:])
        return pieces

    def pre_process(self, state, ref_path, cur_index, ref_speed, **kwargs):
        """Nominal rollout with the previous controls + arc-length stepped reference
        (mpc.py:251-291).  Returns (3 x (T+1) nominal states, list of T+1 reference points,
        index of the closest waypoint)."""
        _, near = self.closest_point(state, ref_path, cur_index, **kwargs)
        cur = state
        ref_pt = ref_path[near]
        refs, preds = [ref_pt], [cur]
        step = ref_speed * self.dt
        for i in range(self.receding):
            vel = self.cur_vel_array[:, i:i + 1]
            if self.dynamics == "acker":
                cur = self.motion_predict_model_acker(cur, vel, self.L, self.dt)
            elif self.dynamics == "diff":
                cur = self.motion_predict_model_diff(cur, vel, self.dt)
            elif self.dynamics == "omni":
                cur = self.motion_predict_model_omni(cur, vel, self.dt)
            preds.append(cur)
            ref_pt, cur_index = self.inter_point(ref_pt, ref_path, cur_index, step)
            # heading of the reference unwrapped around the predicted heading (:285-286);
            # in place, as the reference does
            ref_pt[2, 0] = cur[2, 0] + wrap_to_pi(ref_pt[2, 0] - cur[2, 0])
            refs.append(ref_pt)
        return np.hstack(preds), refs, near

    def motion_predict_model_acker(self, car_state, vel, wheel_base, sample_time):
        assert car_state.shape == (3, 1) and vel.shape == (2, 1)
        th, v, psi = car_state[2, 0], vel[0, 0], vel[1, 0]
        rate = np.array([[v * cos(th)], [v * sin(th)], [v * tan(psi) / wheel_base]])
        return car_state + rate * sample_time

    def motion_predict_model_diff(self, robot_state, vel, sample_time):
        assert robot_state.shape == (3, 1) and vel.shape == (2, 1)
        th, v, w = robot_state[2, 0], vel[0, 0], vel[1, 0]
        return robot_state + np.array([[v * cos(th)], [v * sin(th)], [w]]) * sample_time

    def motion_predict_model_omni(self, robot_state, vel, sample_time):
        assert robot_state.shape[0] >= 2 and vel.shape == (2, 1)
        sp, hd = vel[0, 0], vel[1, 0]
        return robot_state + sample_time * np.array([[sp * cos(hd)], [sp * sin(hd)], [0]])

    def closest_point(self, state, ref_path, start_ind, threshold=0.1, ind_range=10, **kwargs):
        """Nearest waypoint among the next `ind_range`, first one closer than `threshold`
        wins (mpc.py:338-353)."""
        best, best_i = inf, start_ind
        for k, wp in enumerate(ref_path[start_ind:start_ind + ind_range]):
            dk = MPC.distance(state[0:2], wp[0:2])
            if dk < best:
                best, best_i = dk, start_ind + k
                if dk < threshold:
                    break
        return best, best_i

    def inter_point(self, traj_point, ref_path, cur_ind, length):
        """Advance `length` along the polyline from traj_point: far intersection of the
        circle of that radius with the first segment it reaches (mpc.py:355-383)."""
        centre = np.squeeze(traj_point[0:2])
        out = np.copy(traj_point)
        while cur_ind + 1 <= len(ref_path) - 1:
            a, c = ref_path[cur_ind], ref_path[cur_ind + 1]
            hit = self.range_cir_seg(centre, length, [np.squeeze(a[0:2]), np.squeeze(c[0:2])])
            if hit is None:
                cur_ind += 1
                continue
            turn = wrap_to_pi(c[2, 0] - a[2, 0])
            out[0:2, 0] = hit[:]
            out[2, 0] = wrap_to_pi(a[2, 0] + turn / 2)
            return out, cur_ind
        end = ref_path[-1]                      # path exhausted: the last waypoint itself
        end[2] = wrap_to_pi(end[2])
        return end, cur_ind

    def range_cir_seg(self, circle, r, segment):
        return seg_circle_exit(circle, r, segment)

    @staticmethod
    def distance(point1, point2):
        return sqrt((point1[0, 0] - point2[0, 0]) ** 2 + (point1[1, 0] - point2[1, 0]) ** 2)

    @staticmethod
    def wraptopi(radian):
        return wrap_to_pi(radian)

    # ------------------------------------------------------------------ misc API
    def get_adjust_parameters(self):
        return self.rda.get_adjust_parameter()

    def no_ref_path(self):
        return len(self.ref_path) == 0

    def reset(self):
        self.cur_vel_array = np.zeros((2, self.receding))
        self.cur_index = 0
        self.curve_index = 0
        self.rda.reset()
