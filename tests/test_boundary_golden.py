"""Host front end (rda_planner_b200/mpc.py) against values produced by EXECUTING the reference's
own numpy helpers (tests/golden/boundary_golden.json, generator: oracle/gen_golden.py)."""
import copy
import json
import os
from collections import namedtuple

import numpy as np
import pytest

from rda_planner_b200.mpc import MPC, rdaobs, wrap_to_pi, seg_circle_exit, polygon_halfspaces, polygon_order

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, 'golden', 'boundary_golden.json')))
PATH = list(np.load(os.path.join(HERE, 'golden', 'path_track_ref.npy'), allow_pickle=True))
car = namedtuple('car', 'G h cone_type wheelbase max_speed max_acce dynamics')


class RecordingSolver:
    """Same stand-in as oracle/gen_golden.py::RecordingSolver."""

    def __init__(self, receding, *a, **k):
        self.T = receding
        self.calls = []

    def iterative_solve(self, nom_s, nom_u, ref_states, ref_speed, obstacle_list, **kwargs):
        self.calls.append({'nom_s': np.array(nom_s), 'ref': np.hstack(ref_states), 'obs': list(obstacle_list)})
        k = len(self.calls)
        u = np.vstack([np.full(self.T, 2.0 + 0.1 * k), 0.05 * np.sin(0.3 * k + np.arange(self.T))])
        return u, {'ref_traj_list': ref_states, 'opt_state_list': []}

    def reset(self):
        pass


def make(dyn, T, path, **kw):
    ct = car(None, None, 'Rpositive', 3.0, [10, 1], [10, 0.5], dyn)
    return MPC(ct, path, receding=T, sample_time=0.1, solver_cls=RecordingSolver, **kw)


def test_polygon_halfspaces_and_order():
    for rec in GOLD['polygons']:
        v = np.array(rec['vertex'])
        A, b = polygon_halfspaces(v.copy())
        np.testing.assert_allclose(A, rec['A'], atol=1e-12)
        np.testing.assert_allclose(b, rec['b'], atol=1e-12)
        ok, order = polygon_order(v)
        assert ok == rec['convex'] and order == rec['order']


def test_obstacle_conversion():
    m = make('acker', 10, [])
    A, b = m.convert_inequal_circle(np.array([[20.], [34.]]), 1.5, np.zeros((2, 1)))
    np.testing.assert_allclose(A, GOLD['circle_static']['A']); np.testing.assert_allclose(b, GOLD['circle_static']['b'])
    A, b = m.convert_inequal_circle(np.array([[20.], [34.]]), 1.5, np.array([[0.5], [-0.2]]))
    np.testing.assert_allclose(np.array(A), GOLD['circle_moving']['A']); np.testing.assert_allclose(np.array(b), GOLD['circle_moving']['b'], atol=1e-12)
    sq = np.array([[0., 1, 1, 0], [0, 0, 1, 1]])
    A, b = m.convert_inequal_polygon(sq, np.array([[1.0], [0.5]]))
    np.testing.assert_allclose(np.array(A), GOLD['polygon_moving']['A'], atol=1e-12)
    np.testing.assert_allclose(np.array(b), GOLD['polygon_moving']['b'], atol=1e-12)


def test_scalar_helpers():
    for x, y in GOLD['wraptopi']:
        assert abs(wrap_to_pi(x) - y) < 1e-14
    for rec in GOLD['range_cir_seg']:
        hit = seg_circle_exit(np.array(rec['c']), rec['r'], [np.array(rec['p0']), np.array(rec['p1'])])
        if rec['hit'] is None:
            assert hit is None
        else:
            np.testing.assert_allclose(hit, rec['hit'], atol=1e-12)


def test_motion_models():
    m = make('acker', 10, [])
    for rec in GOLD['models']:
        s, u = np.array(rec['s']), np.array(rec['u'])
        np.testing.assert_allclose(m.motion_predict_model_acker(s, u, 3.0, 0.1), rec['acker'], atol=1e-13)
        np.testing.assert_allclose(m.motion_predict_model_diff(s, u, 0.1), rec['diff'], atol=1e-13)
        np.testing.assert_allclose(m.motion_predict_model_omni(s, u, 0.1), rec['omni'], atol=1e-13)


def test_pre_process():
    for rec in GOLD['pre_process']:
        m = make(rec['dynamics'], rec['T'], copy.deepcopy(PATH))
        m.cur_vel_array = np.array(rec['vel'])
        s_pre, refs, idx = m.pre_process(np.array(rec['state']), m.ref_path, rec['index'], 4.0)
        np.testing.assert_allclose(s_pre, rec['state_pre'], atol=1e-12)
        np.testing.assert_allclose(np.hstack(refs), rec['ref'], atol=1e-12)
        assert idx == rec['new_index']


def test_control_loop_with_recording_solver():
    Obs = namedtuple('Obs', 'center radius vertex cone_type velocity')
    obs = [Obs(np.array([[20.], [34.]]), 1.5, None, 'norm2', np.zeros((2, 1))),
           Obs(np.array([[12.], [40.]]), 1.0, None, 'norm2', np.array([[0.3], [0.1]])),
           Obs(None, None, np.array([[31., 33, 33, 31], [28, 28, 24, 24]]), 'Rpositive', np.zeros((2, 1))),
           Obs(None, None, np.array([[11., 12, 12, 11], [44, 44, 45, 45]]), 'Rpositive', np.array([[0.0], [-0.5]]))]
    m = make('acker', 10, copy.deepcopy(PATH))
    for rec in GOLD['control']:
        u, info = m.control(np.array(rec['state']), 4.0, obs)
        call = m.rda.calls[-1]
        np.testing.assert_allclose(u, rec['u'], atol=1e-13)
        assert info['arrive'] == rec['arrive'] and m.cur_index == rec['cur_index']
        np.testing.assert_allclose(call['nom_s'], rec['nom_s'], atol=1e-12)
        np.testing.assert_allclose(call['ref'], rec['ref'], atol=1e-12)
        for o, A0, b0 in zip(call['obs'], rec['obs_A0'], rec['obs_b0']):
            A = o.A[0] if isinstance(o.A, list) else o.A
            b = o.b[0] if isinstance(o.b, list) else o.b
            np.testing.assert_allclose(A, A0, atol=1e-12); np.testing.assert_allclose(b, b0, atol=1e-12)


def test_arrive_and_split():
    rec = GOLD['arrive']
    m = make('diff', 10, copy.deepcopy(PATH))
    m.cur_index = rec['start_index']
    u, info = m.control(np.array(rec['state']), 4.0, [])
    np.testing.assert_allclose(u, rec['u']); assert info['arrive'] == rec['arrive'] and m.cur_index == rec['cur_index']
    gp = [np.array([[float(i)], [0.0], [0.0], [1.0 if i < 5 else -1.0]]) for i in range(9)]
    m = make('acker', 10, gp, enable_reverse=True)
    assert [len(c) for c in m.curve_list] == GOLD['split_path']
