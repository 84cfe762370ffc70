"""Front-end kernels (rda_frontend.cu) through the C ABI: against the g++ build of the same cores, and the
closed loop of BatchedMPC against the host front end (rda_planner_b200/mpc.py) driving the same solver."""
import copy
import os
from collections import namedtuple

import numpy as np
import pytest
import torch

import shim
from rda_planner_b200.scenarios import rectangle_robot

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
PATH = list(np.load(os.path.join(HERE, 'golden', 'path_track_ref.npy'), allow_pickle=True))
PATH_ARR = np.stack([np.asarray(p, float).reshape(-1)[:3] for p in PATH])
Obs = namedtuple('Obs', 'center radius vertex cone_type velocity')


def _random_obstacles(rng, count):
    obs = []
    for j in range(count):
        c = PATH_ARR[int(rng.integers(0, len(PATH_ARR))), :2].reshape(2, 1) + rng.uniform(2.5, 6.0, (2, 1)) * rng.choice([-1, 1], (2, 1))
        vel = rng.uniform(-0.5, 0.5, (2, 1)) if j % 2 else np.zeros((2, 1))
        if j % 3 == 0:
            obs.append(Obs(c, float(rng.uniform(0.3, 1.0)), None, 'norm2', vel))
        else:
            n = int(rng.integers(3, 5))
            ang = np.sort(rng.uniform(0, 2 * np.pi, n))
            if j % 2:
                ang = ang[::-1]
            obs.append(Obs(None, None, c + rng.uniform(0.5, 1.2) * np.vstack([np.cos(ang), np.sin(ang)]), 'Rpositive', vel))
    return obs


@pytest.mark.parametrize('dyn', ['acker', 'diff', 'omni'])
def test_pre_process_kernel_matches_cpu_core(dyn):
    from rda_planner_b200.frontend import pre_process_batch, path_tensor
    rng = np.random.default_rng(3)
    B, T = 96, 15
    idx = rng.integers(0, len(PATH_ARR), B)
    idx[:8] = len(PATH_ARR) - 1 - rng.integers(0, 5, 8)             # exhausted tail
    start = np.maximum(0, idx - rng.integers(0, 8, B)).astype(np.int32)
    state = (PATH_ARR[idx] + rng.normal(0, [0.3, 0.3, 0.2], (B, 3))).astype(np.float32)
    vel = np.stack([rng.uniform(1, 5, (B, T)), rng.uniform(-0.3, 0.3, (B, T))], 1).astype(np.float32)
    speed = rng.uniform(2, 5, B).astype(np.float32)
    dev = torch.device('cuda:0')
    nom, ref, near = pre_process_batch(torch.as_tensor(state, device=dev), torch.as_tensor(vel, device=dev),
                                       torch.as_tensor(speed, device=dev), path_tensor(PATH, dev),
                                       torch.as_tensor(start, device=dev), dyn, 0.1, 3.0, T)
    nom, ref, near = nom.cpu().numpy(), ref.cpu().numpy(), near.cpu().numpy()
    for b in range(B):
        n1, r1, k1 = shim.pre_process(dyn, T, 0.1, 3.0, state[b], vel[b], float(speed[b]), PATH_ARR, int(start[b]))
        assert near[b] == k1
        np.testing.assert_allclose(nom[b], n1, atol=2e-5)
        np.testing.assert_allclose(ref[b], r1, atol=2e-5)


@pytest.mark.parametrize('tv,order', [(False, True), (True, True), (True, False)])
def test_convert_obstacles_kernel_matches_cpu_core(tv, order):
    from rda_planner_b200.frontend import convert_obstacles_batch, pack_shapes, shapes_to_device
    rng = np.random.default_rng(11)
    B, T, N, E, M = 33, 10, 6, 5, 16
    lists = [_random_obstacles(rng, int(c)) for c in rng.integers(0, 14, B)]
    lists[0] = []
    state = (PATH_ARR[rng.integers(0, len(PATH_ARR), B)] + rng.normal(0, 0.3, (B, 3))).astype(np.float32)
    shapes = pack_shapes(lists, M)
    dev = torch.device('cuda:0')
    A, b, kind, count = convert_obstacles_batch(shapes_to_device(shapes, dev), torch.as_tensor(state, device=dev), N, T, E, 0.1, tv, order)
    A, b, kind, count = A.cpu().numpy(), b.cpu().numpy(), kind.cpu().numpy(), count.cpu().numpy()
    for i in range(B):
        one = {k: v[i] for k, v in shapes.items()}
        A1, b1, k1, c1 = shim.convert_obstacles(one, N, T, E, 0.1, tv, order, state[i])
        assert count[i] == c1 == len(lists[i])
        assert list(kind[i]) == list(k1)
        np.testing.assert_array_equal(A[i], A1)
        np.testing.assert_allclose(b[i], b1, rtol=1e-6, atol=1e-6)


def test_batched_mpc_closed_loop_matches_host_front_end():
    """Three robots on the path_track reference with static and moving obstacles, 4 control steps:
    BatchedMPC (everything on the device) against mpc.MPC (host front end) per robot, same solver."""
    from rda_planner_b200.frontend import BatchedMPC, pack_shapes, shapes_to_device
    from rda_planner_b200.mpc import MPC
    car = rectangle_robot()
    T, N, E, steps = 10, 4, 4, 4
    obs = [Obs(np.array([[20.], [34.]]), 1.5, None, 'norm2', np.zeros((2, 1))),
           Obs(np.array([[12.], [40.]]), 1.0, None, 'norm2', np.array([[0.3], [0.1]])),
           Obs(None, None, np.array([[31., 33, 33, 31], [28, 28, 24, 24]]), 'Rpositive', np.zeros((2, 1))),
           Obs(None, None, np.array([[11., 12, 12, 11], [44, 44, 45, 45]]), 'Rpositive', np.array([[0.0], [-0.5]]))]
    starts = [0, 40, len(PATH) - 12]
    B = len(starts)
    states = np.stack([PATH_ARR[i] + np.array([0.2, -0.1, 0.05]) for i in starts]).astype(np.float32)
    kw = dict(receding=T, sample_time=0.1, iter_num=3, max_edge_num=E, max_obs_num=N, iter_threshold=0.0)
    bm = BatchedMPC(car, PATH, B, **kw)
    bm.cur_index[:] = torch.as_tensor(starts, dtype=torch.int32)
    hosts = []
    for i in starts:
        m = MPC(car, copy.deepcopy(PATH), time_print=False, **kw)
        m.cur_index = i
        hosts.append(m)
    dev_state = torch.as_tensor(states, device='cuda:0')
    host_state = [states[i].astype(float).reshape(3, 1) for i in range(B)]
    shapes = shapes_to_device(pack_shapes([obs] * B, 8), 'cuda:0')
    for k in range(steps):
        u0, info = bm.control(dev_state, 4.0, shapes, time_varying=True)
        u0 = u0.cpu().numpy()
        arrive = info['arrive'].cpu().numpy()
        for i, m in enumerate(hosts):
            uh, ih = m.control(host_state[i], 4.0, obs)
            assert bool(arrive[i]) == ih['arrive']
            assert int(info['cur_index'][i]) == m.cur_index
            np.testing.assert_allclose(u0[i], uh[:, 0], atol=2e-3)
            s = host_state[i]
            host_state[i] = s + 0.1 * np.array([[uh[0, 0] * np.cos(s[2, 0])], [uh[0, 0] * np.sin(s[2, 0])], [uh[0, 0] * np.tan(uh[1, 0]) / 3.0]])
        bm.advance(dev_state)
        np.testing.assert_allclose(dev_state.cpu().numpy(), np.hstack(host_state).T, atol=2e-3)


# ---- the kernels against the values produced by EXECUTING the reference's own helpers (no g++ twin in between) ----
import json
GOLD = json.load(open(os.path.join(HERE, 'golden', 'boundary_golden.json')))


def test_pre_process_kernel_matches_reference_goldens():
    """k_pre_process vs MPC.pre_process of the reference (mpc.py:251-291) run by oracle/gen_golden.py."""
    from rda_planner_b200.frontend import pre_process_batch, path_tensor
    dev = torch.device('cuda:0')
    by_cfg = {}
    for rec in GOLD['pre_process']:
        by_cfg.setdefault((rec['dynamics'], rec['T']), []).append(rec)
    assert by_cfg
    for (dyn, T), recs in by_cfg.items():
        state = torch.as_tensor(np.array([np.ravel(r['state'])[:3] for r in recs], np.float32), device=dev)
        vel = torch.as_tensor(np.array([r['vel'] for r in recs], np.float32), device=dev)
        start = torch.as_tensor(np.array([r['index'] for r in recs], np.int32), device=dev)
        speed = torch.full((len(recs),), 4.0, dtype=torch.float32, device=dev)
        nom, ref, near = pre_process_batch(state, vel, speed, path_tensor(PATH, dev), start, dyn, 0.1, 3.0, T)
        for i, r in enumerate(recs):
            assert int(near[i]) == r['new_index']
            np.testing.assert_allclose(nom[i].cpu().numpy(), r['state_pre'], atol=3e-5)
            np.testing.assert_allclose(ref[i].cpu().numpy(), r['ref'], atol=3e-5)


def test_convert_obstacles_kernel_matches_reference_goldens():
    """k_convert_obstacles vs convert_inequal_circle / convert_inequal_polygon / gen_inequal_global of the reference
    (mpc.py:440-510): moving disc, moving square, and the CW / CCW polygons of the reference run."""
    from rda_planner_b200.frontend import convert_obstacles_batch, pack_shapes, shapes_to_device
    from rda_planner_b200 import _cabi
    dev = torch.device('cuda:0')
    sq = np.array([[0., 1, 1, 0], [0, 0, 1, 1]])
    lst = [Obs(np.array([[20.], [34.]]), 1.5, None, 'norm2', np.array([[0.5], [-0.2]])),
           Obs(None, None, sq, 'Rpositive', np.array([[1.0], [0.5]]))]
    shapes = shapes_to_device(pack_shapes([lst], 4), dev)
    A, b, kind, cnt = convert_obstacles_batch(shapes, torch.zeros((1, 3), device=dev), 2, 10, 4, 0.1, True, False)
    A, b, kind = A[0].cpu().numpy(), b[0].cpu().numpy(), kind[0].cpu().numpy()
    assert int(cnt[0]) == 2 and list(kind) == [_cabi.OBS_CIRCLE, _cabi.OBS_POLYGON]
    np.testing.assert_allclose(A[0, :, :3], GOLD['circle_moving']['A'], atol=1e-6)
    np.testing.assert_allclose(b[0, :, :3], np.array(GOLD['circle_moving']['b'])[:, :, 0], atol=1e-5)
    np.testing.assert_allclose(A[1], GOLD['polygon_moving']['A'], atol=1e-6)
    np.testing.assert_allclose(b[1], np.array(GOLD['polygon_moving']['b'])[:, :, 0], atol=1e-5)
    n_checked = 0
    for rec in GOLD['polygons']:
        v = np.array(rec['vertex'])
        if v.shape[1] > 8:
            continue
        shapes = shapes_to_device(pack_shapes([[Obs(None, None, v, 'Rpositive', np.zeros((2, 1)))]], 1), dev)
        A, b, kind, cnt = convert_obstacles_batch(shapes, torch.zeros((1, 3), device=dev), 1, 5, 8, 0.1, False, False)
        n = v.shape[1]
        np.testing.assert_allclose(A[0, 0, 0, :n].cpu().numpy(), rec['A'], atol=1e-5)
        np.testing.assert_allclose(b[0, 0, 0, :n].cpu().numpy(), np.array(rec['b'])[:, 0], atol=1e-4)
        n_checked += 1
    assert n_checked >= 2


@pytest.mark.parametrize('nobs', [0, 3])
def test_arrive_rule_and_free_space_match_reference_golden(nobs):
    """End of the path, no obstacles (mpc.py:166-187 run by oracle/gen_golden.py): controls zeroed, arrive flag, index.
    max_obs_num = 0 exercises the N == 0 paths of rda_create / k_finalize."""
    from rda_planner_b200.frontend import BatchedMPC
    g = GOLD['arrive']
    bm = BatchedMPC(rectangle_robot(dynamics='diff'), PATH, 1, receding=10, sample_time=0.1, iter_num=2, max_obs_num=nobs)
    bm.cur_index[:] = g['start_index']
    u0, info = bm.control(np.ravel(g['state'])[:3][None].astype(np.float32), 4.0, None)
    assert bool(info['arrive'][0]) == g['arrive'] and int(info['cur_index'][0]) == g['cur_index']
    np.testing.assert_allclose(u0[0].cpu().numpy(), np.ravel(g['u']), atol=1e-6)
    assert int(info['status'][0]) & 6 == 0


def test_zero_obstacle_slots_single_instance_api():
    """RDA_solver(max_obs_num=0).iterative_solve: pure tracking su-QP, no cell kernels, no NULL obs_count read."""
    from rda_planner_b200.rda_solver import RDA_solver
    from oracle.rda_oracle import OracleRDA
    from rda_planner_b200.scenarios import make_instance
    T = 8
    car = rectangle_robot()
    inst = make_instance(5, T=T, N=1, E=4)
    ref = [inst['ref'][:, t:t + 1] for t in range(T + 1)]
    g = RDA_solver(T, car, max_edge_num=4, max_obs_num=0, iter_num=3, iter_threshold=0.0, time_print=False)
    o = OracleRDA(T, car, max_edge_num=4, max_obs_num=0, iter_num=3, iter_threshold=0.0)
    ug, ig = g.iterative_solve(inst['nom_s'], inst['nom_u'], ref, 4.0, [])
    uo, io = o.iterative_solve(inst['nom_s'], inst['nom_u'], ref, 4.0, [])
    assert ig['status'] & 6 == 0
    np.testing.assert_allclose(ug, uo, atol=1e-3)
    np.testing.assert_allclose(np.hstack(ig['opt_state_list']), np.hstack(io['opt_state_list']), atol=1e-3)


def test_batched_mpc_gear_changes_match_host_front_end():
    """enable_reverse (mpc.py:139-144, :166-183, split_path :232-249): a path that drives 5 m forward, 4 m back and
    forward again, gear flag in the 4th row.  Two robots at different places of it, closed loop over enough steps to
    switch curves; BatchedMPC against mpc.MPC per robot (same solver): controls, arrive flag, curve and waypoint index."""
    from rda_planner_b200.frontend import BatchedMPC
    from rda_planner_b200.mpc import MPC
    car = rectangle_robot()
    pts = []
    for i in range(26):                                   # forward along +x
        pts.append(np.array([[0.2 * i], [0.0], [0.0], [1.0]]))
    for i in range(1, 21):                                # reverse along -x (heading still 0)
        pts.append(np.array([[5.0 - 0.2 * i], [0.0], [0.0], [-1.0]]))
    for i in range(1, 16):                                # forward again
        pts.append(np.array([[1.0 + 0.2 * i], [0.0], [0.0], [1.0]]))
    T, steps = 8, 30
    kw = dict(receding=T, sample_time=0.1, iter_num=2, max_edge_num=4, max_obs_num=2, iter_threshold=0.0, enable_reverse=True)
    starts = [(0, 0), (1, 2)]                             # (curve, waypoint index inside the curve)
    B = len(starts)
    bm = BatchedMPC(car, pts, B, **kw)
    assert bm.n_curves == 3 and bm.curve_gear.cpu().tolist() == [1.0, -1.0, 1.0]
    hosts, host_state, st0 = [], [], []
    for c, i in starts:
        m = MPC(car, copy.deepcopy(pts), time_print=False, **kw)
        assert [len(x) for x in m.curve_list] == [26, 20, 15]
        m.curve_index, m.cur_index = c, i
        hosts.append(m)
        wp = np.asarray(m.curve_list[c][i], float).reshape(-1)[:3]
        st0.append(wp + np.array([0.05, 0.03, 0.01]))
        host_state.append(st0[-1].reshape(3, 1).copy())
    bm.curve_index[:] = torch.as_tensor([c for c, _ in starts], dtype=torch.int32)
    bm.cur_index[:] = torch.as_tensor([i for _, i in starts], dtype=torch.int32)
    dev_state = torch.as_tensor(np.stack(st0).astype(np.float32), device='cuda:0')
    switched = 0
    for k in range(steps):
        u0, info = bm.control(dev_state, 2.0, None)
        u0 = u0.cpu().numpy()
        for i, m in enumerate(hosts):
            before = m.curve_index
            uh, ih = m.control(host_state[i], 2.0, [])
            switched += int(m.curve_index != before)
            assert bool(info['arrive'][i]) == ih['arrive'], (k, i)
            assert int(info['curve_index'][i]) == min(m.curve_index, 2), (k, i)
            assert int(info['cur_index'][i]) == m.cur_index, (k, i)
            np.testing.assert_allclose(u0[i], uh[:, 0], atol=3e-3)
            s = host_state[i]
            host_state[i] = s + 0.1 * np.array([[uh[0, 0] * np.cos(s[2, 0])], [uh[0, 0] * np.sin(s[2, 0])], [uh[0, 0] * np.tan(uh[1, 0]) / 3.0]])
            if m.curve_index >= len(m.curve_list):
                m.curve_index = len(m.curve_list) - 1      # the reference raises IndexError on its next call; stay arrived
                m.cur_index = len(m.curve_list[-1]) - 1
        bm.advance(dev_state)
        np.testing.assert_allclose(dev_state.cpu().numpy(), np.hstack(host_state).T, atol=3e-3)
    assert switched >= 2
