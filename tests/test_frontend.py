"""Front-end cores (rda_planner_b200/csrc/frontend.cuh, g++ build) against
 (1) values produced by EXECUTING the reference's numpy helpers (tests/golden/boundary_golden.json),
 (2) the host front end rda_planner_b200/mpc.py + pack_obstacles (itself pinned to the same goldens).
float32 inputs/outputs, float64 arithmetic: stated tolerance 2e-5 absolute on coordinates of O(50 m)."""
import copy
import json
import os
from collections import namedtuple

import numpy as np
import pytest

import shim
from rda_planner_b200.mpc import MPC
from rda_planner_b200 import _cabi

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, 'golden', 'boundary_golden.json')))
PATH = list(np.load(os.path.join(HERE, 'golden', 'path_track_ref.npy'), allow_pickle=True))
PATH_ARR = np.stack([np.asarray(p, float).reshape(-1)[:3] for p in PATH])
car = namedtuple('car', 'G h cone_type wheelbase max_speed max_acce dynamics')
Obs = namedtuple('Obs', 'center radius vertex cone_type velocity')
TOL = 2e-5


class _NoSolver:
    def __init__(self, *a, **k):
        pass


def host_mpc(dyn, T, path):
    return MPC(car(None, None, 'Rpositive', 3.0, [10, 1], [10, 0.5], dyn), path, receding=T, sample_time=0.1,
               solver_cls=_NoSolver)


def test_pre_process_matches_reference_goldens():
    for rec in GOLD['pre_process']:
        nom, ref, near = shim.pre_process(rec['dynamics'], rec['T'], 0.1, 3.0, np.array(rec['state']), np.array(rec['vel']),
                                          4.0 * 0.1 / 0.1, PATH_ARR, rec['index'])
        np.testing.assert_allclose(nom, rec['state_pre'], atol=TOL)
        np.testing.assert_allclose(ref, rec['ref'], atol=TOL)
        assert near == rec['new_index']


@pytest.mark.parametrize('dyn', ['acker', 'diff', 'omni'])
def test_pre_process_matches_host_front_end_along_the_path(dyn):
    """Random poses along the whole path, including the exhausted tail (aliased last waypoint)."""
    rng = np.random.default_rng(7)
    T = 12
    for trial in range(40):
        idx = int(rng.integers(0, len(PATH))) if trial % 4 else len(PATH) - 1 - int(rng.integers(0, 6))
        start = max(0, idx - int(rng.integers(0, 8)))
        wp = PATH_ARR[idx]
        state = np.array([[wp[0] + rng.normal(0, 0.3)], [wp[1] + rng.normal(0, 0.3)], [wp[2] + rng.normal(0, 0.2)]])
        state = state.astype(np.float32).astype(float)
        vel = np.vstack([rng.uniform(1, 5, T), rng.uniform(-0.3, 0.3, T)]).astype(np.float32).astype(float)
        m = host_mpc(dyn, T, copy.deepcopy(PATH))
        m.cur_vel_array = vel
        s_pre, refs, near_h = m.pre_process(state, m.ref_path, start, 4.0)
        nom, ref, near = shim.pre_process(dyn, T, 0.1, 3.0, state, vel, 4.0, PATH_ARR, start)
        assert near == near_h
        np.testing.assert_allclose(nom, s_pre, atol=TOL)
        np.testing.assert_allclose(ref, np.hstack([r[0:3] for r in refs]), atol=TOL)


def _shapes_from(obs_list, M):
    from rda_planner_b200.frontend import pack_shapes
    s = pack_shapes([obs_list], M)
    return {k: v[0] for k, v in s.items()}


def test_obstacle_rows_match_reference_goldens():
    sq = np.array([[0., 1, 1, 0], [0, 0, 1, 1]])
    shapes = _shapes_from([Obs(np.array([[20.], [34.]]), 1.5, None, 'norm2', np.array([[0.5], [-0.2]])),
                           Obs(None, None, sq, 'Rpositive', np.array([[1.0], [0.5]]))], 4)
    A, b, kind, cnt = shim.convert_obstacles(shapes, 2, 10, 4, 0.1, True, False, np.zeros(3))
    assert cnt == 2 and list(kind) == [_cabi.OBS_CIRCLE, _cabi.OBS_POLYGON]
    np.testing.assert_allclose(A[0, :, :3], GOLD['circle_moving']['A'], atol=1e-6)
    np.testing.assert_allclose(b[0, :, :3], np.array(GOLD['circle_moving']['b'])[:, :, 0], atol=1e-5)
    assert np.all(A[0, :, 3] == 0) and np.all(b[0, :, 3] == 0)
    np.testing.assert_allclose(A[1], GOLD['polygon_moving']['A'], atol=1e-6)
    np.testing.assert_allclose(b[1], np.array(GOLD['polygon_moving']['b'])[:, :, 0], atol=1e-5)
    for rec in GOLD['polygons']:                      # CW / CCW / non-convex inputs of the reference run
        v = np.array(rec['vertex'])
        if v.shape[1] > 8:
            continue
        shapes = _shapes_from([Obs(None, None, v, 'Rpositive', np.zeros((2, 1)))], 1)
        A, b, kind, cnt = shim.convert_obstacles(shapes, 1, 5, 8, 0.1, False, False, np.zeros(3))
        n = v.shape[1]
        np.testing.assert_allclose(A[0, 0, :n], rec['A'], atol=1e-5)
        np.testing.assert_allclose(b[0, 0, :n], np.array(rec['b'])[:, 0], atol=1e-4)


@pytest.mark.parametrize('order', [False, True])
@pytest.mark.parametrize('count', [0, 1, 3, 9])
def test_convert_obstacles_matches_host_front_end(order, count):
    """Sorting by distance, truncation to N, padding by repetition, zero rows, moving shapes."""
    from rda_planner_b200.rda_solver import pack_obstacles
    rng = np.random.default_rng(100 + count)
    T, N, E, M = 8, 5, 5, 12
    state = np.array([[10.0], [5.0], [0.3]])
    obs = []
    for j in range(count):
        c = rng.uniform(0, 20, (2, 1))
        vel = rng.uniform(-1, 1, (2, 1)) if j % 2 else np.zeros((2, 1))
        if j % 3 == 0:
            obs.append(Obs(c, float(rng.uniform(0.3, 1.5)), None, 'norm2', vel))
        else:
            n = int(rng.integers(3, 6))
            ang = np.sort(rng.uniform(0, 2 * np.pi, n))
            if j % 2:
                ang = ang[::-1]                        # clockwise input
            v = c + rng.uniform(0.5, 2.0) * np.vstack([np.cos(ang), np.sin(ang)])
            obs.append(Obs(None, None, v, 'Rpositive', vel))
    obs = [o._replace(center=None if o.center is None else o.center.astype(np.float32).astype(float),
                      vertex=None if o.vertex is None else o.vertex.astype(np.float32).astype(float),
                      velocity=o.velocity.astype(np.float32).astype(float)) for o in obs]
    m = host_mpc('acker', T, [])
    m.state = state
    rda_obs = m.convert_rda_obstacle(obs, state, order)
    tv = any(isinstance(o.A, list) for o in rda_obs[:N])
    shapes = _shapes_from(obs, M)
    A, b, kind, cnt = shim.convert_obstacles(shapes, N, T, E, 0.1, tv, order, state)
    assert cnt == count
    if count == 0:
        assert not A.any() and not b.any()
        return
    Ah, bh, kh, ch, tvh = pack_obstacles(list(rda_obs), T, N, E)
    assert tvh == tv and ch == count
    assert list(kind) == list(kh)
    np.testing.assert_allclose(A, Ah, atol=1e-5)
    np.testing.assert_allclose(b, bh, atol=1e-4)


# ---- size-independent properties of the front-end cores (hypothesis) ------------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None, derandomize=True, database=None)
@given(n=st.integers(3, 8), cx=st.floats(-50, 50), cy=st.floats(-50, 50), rad=st.floats(0.3, 3.0),
       clockwise=st.booleans(), vx=st.floats(-1, 1), vy=st.floats(-1, 1), seed=st.integers(0, 10 ** 6))
def test_polygon_rows_contain_the_shape_and_follow_its_motion(n, cx, cy, rad, clockwise, vx, vy, seed):
    """For any convex polygon, either orientation: every vertex satisfies A v <= b (tight on its two rows),
    the centroid is strictly inside, rows turn counter-clockwise, and the copy at stage t is the shape moved
    by velocity * t * dt (|velocity| > 0.01) or the same shape (slower)."""
    rng = np.random.default_rng(seed)
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    if np.min(np.diff(np.concatenate([ang, [ang[0] + 2 * np.pi]]))) < 0.15:
        ang = np.linspace(0, 2 * np.pi, n, endpoint=False) + rng.uniform(0, 1)
    v = np.array([[cx], [cy]]) + rad * np.vstack([np.cos(ang), np.sin(ang)])
    if clockwise:
        v = v[:, ::-1]
    v = v.astype(np.float32).astype(float)
    vel = np.array([[vx], [vy]], np.float32).astype(float)
    T, dt = 6, 0.1
    shapes = _shapes_from([Obs(None, None, v, 'Rpositive', vel)], 1)
    A, b, kind, cnt = shim.convert_obstacles(shapes, 1, T, 8, dt, True, False, np.zeros(3))
    moving = np.hypot(*vel[:, 0]) > 0.01
    for t in (0, T):
        off = vel * (t * dt) if moving else 0.0
        vt = v + off
        At, bt = A[0, t, :n].astype(float), b[0, t, :n].astype(float)
        assert not A[0, t, n:].any() and not b[0, t, n:].any()
        viol = At @ vt - bt[:, None]
        scale = 1 + np.abs(bt).max()
        assert viol.max() < 2e-5 * scale
        assert np.all(At @ vt.mean(1) - bt < 0)
        nxt = np.roll(At, -1, axis=0)
        assert np.all(At[:, 0] * nxt[:, 1] - At[:, 1] * nxt[:, 0] > 0)          # rows counter-clockwise
        assert np.sort(np.abs(viol), axis=1)[:, :2].max() < 2e-5 * scale         # each row tight on two vertices


@settings(max_examples=40, deadline=None, derandomize=True, database=None)
@given(idx=st.integers(0, len(PATH_ARR) - 1), back=st.integers(0, 8), speed=st.floats(1.0, 6.0),
       dyn=st.sampled_from(['acker', 'diff', 'omni']), seed=st.integers(0, 10 ** 6))
def test_pre_process_properties(idx, back, speed, dyn, seed):
    """The nominal columns are the model rolled out with the given controls; every reference point lies on
    the path polyline; consecutive reference points are one arc step ref_speed * dt apart until the path
    is exhausted, after which they sit on the last waypoint; headings are unwrapped around the prediction."""
    from rda_planner_b200.scenarios import rollout
    rng = np.random.default_rng(seed)
    T, dt = 12, 0.1
    start = max(0, idx - back)
    state = (PATH_ARR[idx] + rng.normal(0, [0.2, 0.2, 0.1])).astype(np.float32)
    vel = np.vstack([rng.uniform(0.5, 5, T), rng.uniform(-0.3, 0.3, T)]).astype(np.float32)
    nom, ref, near = shim.pre_process(dyn, T, dt, 3.0, state, vel, float(speed), PATH_ARR, start)
    assert start <= near < min(start + 10, len(PATH_ARR))
    np.testing.assert_allclose(nom, rollout(state.astype(float), vel.astype(float), dt, 3.0, dyn), atol=2e-5)
    P32 = PATH_ARR.astype(np.float32).astype(float)
    seg_a, seg_b = P32[:-1, :2], P32[1:, :2]
    for t in range(1, T + 1):
        p = ref[:2, t].astype(float)
        d = seg_b - seg_a
        tt = np.clip(np.einsum('ij,ij->i', p - seg_a, d) / np.maximum(np.einsum('ij,ij->i', d, d), 1e-30), 0, 1)
        assert np.min(np.linalg.norm(seg_a + tt[:, None] * d - p, axis=1)) < 2e-4      # on the polyline
        stepped = np.linalg.norm(ref[:2, t].astype(float) - ref[:2, t - 1].astype(float))
        at_end = np.linalg.norm(p - P32[-1, :2]) < 1e-6
        if t >= 2:
            assert at_end or abs(stepped - speed * dt) < 2e-4
        assert abs(float(ref[2, t]) - float(nom[2, t])) <= np.pi + 1e-4                 # unwrapped around the prediction
