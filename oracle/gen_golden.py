"""TEST INFRASTRUCTURE — generate tests/golden/boundary_golden.json by EXECUTING the reference.

The reference's solver stack (cvxpy / ECOS / pathos) is absent from this image, but its
pure-numpy helpers run once `cvxpy` and `pathos` are stubbed.  This script
imports /root/reference/RDA_planner (read-only; nothing is copied), calls

  MPC.gen_inequal_global / is_convex_and_ordered / convert_inequal_circle / convert_inequal_polygon
  (mpc.py:440-549), MPC.pre_process / closest_point / inter_point / range_cir_seg / wraptopi
  (mpc.py:251-438), MPC.motion_predict_model_* (mpc.py:293-336), MPC.control with a recording
  stand-in for `self.rda` (mpc.py:127-187), MPC.convert_rda_obstacle (mpc.py:189-218) and
  RDA_solver.linear_{ackermann,diff,omni}_model (rda_solver.py:949-994)

on seeded inputs and stores inputs + outputs as JSON.  Run it in the build container only
(/root/reference does not exist on the GPU box); the JSON is committed.

    python oracle/gen_golden.py
"""
import copy
import json
import os
import sys
from collections import namedtuple
from unittest import mock

import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'boundary_golden.json')


def load_reference():
    for name in ('cvxpy', 'pathos', 'pathos.multiprocessing'):
        sys.modules[name] = mock.MagicMock()
    sys.path.insert(0, REF)
    for k in list(sys.modules):
        if k == 'RDA_planner' or k.startswith('RDA_planner.'):
            del sys.modules[k]
    import RDA_planner.mpc as ref_mpc
    import RDA_planner.rda_solver as ref_solver
    assert ref_mpc.__file__.startswith(REF), ref_mpc.__file__
    return ref_mpc, ref_solver


def tolist(x):
    if isinstance(x, (list, tuple)):
        return [tolist(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    return x


class RecordingSolver:
    """Stand-in for RDA_solver: returns a fixed control sequence and records its inputs."""

    def __init__(self, T):
        self.T = T
        self.calls = []

    def iterative_solve(self, nom_s, nom_u, ref_states, ref_speed, obstacle_list, **kwargs):
        self.calls.append({'nom_s': np.array(nom_s), 'nom_u': np.array(nom_u),
                           'ref': np.hstack(ref_states), 'ref_speed': ref_speed,
                           'n_obs': len(obstacle_list),
                           'obs_A0': [np.array(o.A[0] if isinstance(o.A, list) else o.A) for o in obstacle_list],
                           'obs_b0': [np.array(o.b[0] if isinstance(o.b, list) else o.b) for o in obstacle_list]})
        k = len(self.calls)
        u = np.vstack([np.full(self.T, 2.0 + 0.1 * k), 0.05 * np.sin(0.3 * k + np.arange(self.T))])
        return u, {'ref_traj_list': ref_states, 'opt_state_list': [nom_s[:, i:i + 1] for i in range(nom_s.shape[1])]}

    def reset(self):
        pass


def make_ref_mpc(ref_mpc, dynamics, T, path, enable_reverse=False):
    car = namedtuple('car', 'G h cone_type wheelbase max_speed max_acce dynamics')
    m = ref_mpc.MPC.__new__(ref_mpc.MPC)
    m.car_tuple = car(None, None, 'Rpositive', 3.0, [10, 1], [10, 0.5], dynamics)
    m.L, m.dynamics, m.receding, m.dt = 3.0, dynamics, T, 0.1
    m.cur_vel_array = np.zeros((2, T))
    m.state = np.zeros((3, 1))
    m.cur_index = 0
    m.ref_path = path
    m.rda = RecordingSolver(T)
    m.enable_reverse = enable_reverse
    m.rda_obstacle = False
    m.obstacle_order = True
    m.goal_index_threshold = 1
    if enable_reverse:
        m.curve_list = m.split_path(m.ref_path)
        m.curve_index = 0
    return m


def main():
    ref_mpc, ref_solver = load_reference()
    rng = np.random.default_rng(20260923)
    G = {}
    m = make_ref_mpc(ref_mpc, 'acker', 10, [])
    # ---- polygons ----
    polys = [np.array([[-0.8, 3.8, 3.8, -0.8], [-0.8, -0.8, 0.8, 0.8]]),
             np.array([[31., 33, 33, 31], [28, 28, 24, 24]]),          # CW (path_track.yaml)
             np.array([[0., 2, 1], [0, 0, 3]]),
             np.array([[0., 1, 1.5, 1, 0, -0.5], [0, 0, 1, 2, 2, 1]])]
    G['polygons'] = []
    for v in polys:
        A, b = m.gen_inequal_global(v.copy())
        ok, order = m.is_convex_and_ordered(v)
        G['polygons'].append({'vertex': tolist(v), 'A': tolist(A), 'b': tolist(b), 'convex': bool(ok), 'order': order})
    # ---- moving obstacles ----
    A, b = m.convert_inequal_circle(np.array([[20.], [34.]]), 1.5, np.zeros((2, 1)))
    G['circle_static'] = {'A': tolist(A), 'b': tolist(b)}
    A, b = m.convert_inequal_circle(np.array([[20.], [34.]]), 1.5, np.array([[0.5], [-0.2]]))
    G['circle_moving'] = {'A': tolist(A), 'b': tolist(b)}
    sq = np.array([[0., 1, 1, 0], [0, 0, 1, 1]])
    A, b = m.convert_inequal_polygon(sq, np.array([[1.0], [0.5]]))
    G['polygon_moving'] = {'A': tolist(A), 'b': tolist(b)}
    # ---- scalar helpers ----
    G['wraptopi'] = [[x, ref_mpc.MPC.wraptopi(x)] for x in (4.0, -4.0, 0.3, 7.0, -9.5)]
    segs = []
    for _ in range(12):
        c = rng.normal(size=2); r = rng.uniform(0.2, 2.0); p0 = rng.normal(size=2); p1 = p0 + rng.normal(size=2)
        hit = m.range_cir_seg(c, r, [p0, p1])
        segs.append({'c': tolist(c), 'r': r, 'p0': tolist(p0), 'p1': tolist(p1), 'hit': None if hit is None else tolist(hit)})
    G['range_cir_seg'] = segs
    # ---- motion models and jacobians ----
    mm = []
    for _ in range(6):
        st = rng.normal(size=(3, 1)); ut = rng.normal(size=(2, 1)) * np.array([[3.0], [0.4]])
        rec = {'s': tolist(st), 'u': tolist(ut)}
        rec['acker'] = tolist(m.motion_predict_model_acker(st, ut, 3.0, 0.1))
        rec['diff'] = tolist(m.motion_predict_model_diff(st, ut, 0.1))
        rec['omni'] = tolist(m.motion_predict_model_omni(st, ut, 0.1))
        S = ref_solver.RDA_solver
        for name, out in (('acker', S.linear_ackermann_model(None, st, ut, 0.1, 3.0)),
                          ('diff', S.linear_diff_model(None, st, ut, 0.1)),
                          ('omni', S.linear_omni_model(None, ut, 0.1))):
            rec['lin_' + name] = [tolist(o) for o in out]
        mm.append(rec)
    G['models'] = mm
    # ---- pre_process on the shipped path fixture ----
    path = list(np.load(os.path.join(REF, 'example/path_track/path_track_ref.npy'), allow_pickle=True))
    pp = []
    for dyn, T in (('acker', 10), ('diff', 20), ('omni', 30)):
        mp = make_ref_mpc(ref_mpc, dyn, T, copy.deepcopy(path))
        mp.cur_vel_array = np.vstack([np.full(T, 2.0), 0.1 * np.cos(np.arange(T))])
        for state, idx in ((np.array([[10.], [42.], [1.57]]), 0), (np.array([[14.5], [41.0], [0.2]]), 3),
                           (np.array([[28.5], [19.0], [2.9]]), 128)):
            mp.ref_path = copy.deepcopy(path)
            s_pre, refs, new_idx = mp.pre_process(state, mp.ref_path, idx, 4.0)
            pp.append({'dynamics': dyn, 'T': T, 'state': tolist(state), 'index': idx, 'vel': tolist(mp.cur_vel_array),
                       'state_pre': tolist(s_pre), 'ref': tolist(np.hstack(refs)), 'new_index': int(new_idx)})
    G['pre_process'] = pp
    # ---- whole control() loop with a recording solver, incl. obstacle conversion / sorting / arrive ----
    Obs = namedtuple('Obs', 'center radius vertex cone_type velocity')
    obs = [Obs(np.array([[20.], [34.]]), 1.5, None, 'norm2', np.zeros((2, 1))),
           Obs(np.array([[12.], [40.]]), 1.0, None, 'norm2', np.array([[0.3], [0.1]])),
           Obs(None, None, np.array([[31., 33, 33, 31], [28, 28, 24, 24]]), 'Rpositive', np.zeros((2, 1))),
           Obs(None, None, np.array([[11., 12, 12, 11], [44, 44, 45, 45]]), 'Rpositive', np.array([[0.0], [-0.5]]))]
    mp = make_ref_mpc(ref_mpc, 'acker', 10, copy.deepcopy(path))
    state = np.array([[10.], [42.], [1.57], [0.0]])
    ctrl = []
    for k in range(6):
        u, info = mp.control(state, 4.0, obs)
        call = mp.rda.calls[-1]
        ctrl.append({'state': tolist(state), 'u': tolist(u), 'arrive': bool(info['arrive']), 'cur_index': int(mp.cur_index),
                     'nom_s': tolist(call['nom_s']), 'ref': tolist(call['ref']), 'obs_A0': tolist(call['obs_A0']),
                     'obs_b0': tolist(call['obs_b0'])})
        th = state[2, 0]
        state = state + 0.1 * np.array([[u[0, 0] * np.cos(th)], [u[0, 0] * np.sin(th)], [u[0, 0] * np.tan(u[1, 0]) / 3.0], [0.0]])
    G['control'] = ctrl
    # ---- arrive logic near the end of the path ----
    mp = make_ref_mpc(ref_mpc, 'diff', 10, copy.deepcopy(path))
    mp.cur_index = 130
    state = np.array([[28.9], [18.2], [3.0]])
    u, info = mp.control(state, 4.0, [])
    G['arrive'] = {'state': tolist(state), 'start_index': 130, 'u': tolist(u), 'arrive': bool(info['arrive']),
                   'cur_index': int(mp.cur_index)}
    # ---- split_path with gear flags ----
    gp = [np.array([[float(i)], [0.0], [0.0], [1.0 if i < 5 else -1.0]]) for i in range(9)]
    mp = make_ref_mpc(ref_mpc, 'acker', 10, gp, enable_reverse=True)
    G['split_path'] = [len(c) for c in mp.curve_list]
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, 'w') as f:
        json.dump(G, f)
    print('wrote', os.path.abspath(OUT), os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
