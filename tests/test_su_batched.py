"""The batched su-QP pipeline (rda_planner_b200/csrc/su_batched.cuh: one kernel per interior point phase,
[index][instance] workspace) against su_solve (su_solver.cuh, one instance at a time) — both compiled for the
host by oracle/cpu_port; the batched kernels run through a serial launch emulation, which is exact because
they use neither shared memory nor intra-block synchronisation.  The -m gpu tests repeat it on the device."""
import ctypes as C
import os

import numpy as np
import pytest

import shim
from rda_planner_b200.scenarios import rectangle_robot, make_instance


def _params(T, N, dyn='acker', acc=1):
    return shim.SuParams(T=T, N=N, dynamics=shim.DYN[dyn], accelerated=acc, dt=0.1, L=3.0,
                         umax=(C.c_float * 2)(10, 1), ab=(C.c_float * 2)(1.0, 0.05), ws=1, wu=1, slack_gain=8,
                         dmin=0.1, dmax=1.0, ro1=200, ro2=1, max_iter=40, mu0=1.0)


def _inputs(nb, T, N, dyn, seed):
    """nb su-QPs with hinge rows that are partly active: unit normals, margins in [-0.5, 3]."""
    rng = np.random.default_rng(seed)
    cur_s = np.zeros((nb, 3, T + 1), np.float32); cur_u = np.zeros((nb, 2, T), np.float32)
    ref_s = np.zeros((nb, 3, T + 1), np.float32); pref = np.zeros((nb, 2, T), np.float32)
    coef = np.zeros((nb, 5, N, T), np.float32); dis = np.ones((nb, T), np.float32)
    vref = np.full(nb, 4.0, np.float32)
    for b in range(nb):
        inst = make_instance(seed * 100 + b, T=T, N=max(N, 1), E=4, dynamics=dyn, lateral=(0.3, 3.5))
        cur_s[b] = inst['nom_s']; cur_u[b] = inst['nom_u']; ref_s[b] = inst['ref']
        pref[b] = inst['nom_s'][0:2, 1:] + rng.normal(0, 0.05, (2, T))
        ang = rng.uniform(-np.pi, np.pi, (N, T))
        coef[b, 0] = np.cos(ang); coef[b, 1] = np.sin(ang)
        coef[b, 2] = rng.uniform(-0.5, 3.0, (N, T))
        coef[b, 3:5] = rng.normal(0, 0.3, (2, N, T))
        dis[b] = rng.uniform(0.1, 1.0, T)
    return cur_s, cur_u, ref_s, pref, coef, dis, vref


def _batched(P, cur_s, cur_u, ref_s, pref, coef, dis, vref, done=None, max_iter=28):
    nb = cur_s.shape[0]
    cs, cu, dd = cur_s.copy(), cur_u.copy(), dis.copy()
    status = np.zeros(nb, np.int32); iters = np.zeros(nb, np.int32); counters = np.zeros(8, np.int32)
    done = np.zeros(nb, np.int32) if done is None else np.ascontiguousarray(done, np.int32)
    f = shim.lib().shim_su_batched
    f.restype = C.c_int
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    left = f(C.byref(P), C.c_int(nb), p(cs), p(cu), p(ref_s), p(pref), p(coef), p(dd), p(vref), p(done), p(status),
             p(iters), p(counters), C.c_int(max_iter))
    assert left == 0
    return cs, cu, dd, status, iters, counters


@pytest.mark.parametrize('dyn,N,acc', [('acker', 6, 1), ('diff', 5, 1), ('omni', 4, 1), ('acker', 4, 0), ('acker', 0, 1)])
def test_batched_equals_single_instance(dyn, N, acc):
    T, nb = 12, 5
    P = _params(T, N, dyn, acc)
    cur_s, cur_u, ref_s, pref, coef, dis, vref = _inputs(nb, T, N, dyn, 7)
    done = np.array([0, 0, 1, 0, 0], np.int32)                       # instance 2 already stopped (ADMM early stop)
    cs, cu, dd, status, iters, counters = _batched(P, cur_s, cur_u, ref_s, pref, coef, dis, vref, done)
    total_it = 0
    for b in range(nb):
        if done[b]:
            assert np.array_equal(cs[b], cur_s[b]) and np.array_equal(cu[b], cur_u[b]) and iters[b] == 0
            continue
        z = np.zeros((max(N, 1), T), np.float32)
        hx, hy, hc, gx, gy = (coef[b, k] if N > 0 else z[:0] for k in range(5))
        s, u, d, st, it = shim.su(P, cur_s[b].astype(float), cur_u[b].astype(float), ref_s[b].astype(float),
                                  float(vref[b]), dis[b].astype(float), hx, hy, hc, gx, gy, pref[b].astype(float))
        assert st == 0 and status[b] == 0 and iters[b] == 1
        total_it += it
        # same arithmetic, same order: identical up to the float32 rounding of the outputs
        np.testing.assert_allclose(cs[b], s.astype(np.float32), atol=2e-6)
        np.testing.assert_allclose(cu[b], u.astype(np.float32), atol=2e-6)
        np.testing.assert_allclose(dd[b], d.astype(np.float32), atol=2e-6)
    assert counters[3] == total_it and counters[4] == nb - 1


def test_iteration_cap_is_reported_like_optimal_inaccurate():
    T, N, nb = 10, 4, 3
    P = _params(T, N)
    args = _inputs(nb, T, N, 'acker', 11)
    cs, cu, dd, status, iters, counters = _batched(P, *args, max_iter=3)
    assert np.all(status == 1) and np.all(np.isfinite(cs)) and counters[3] == 3 * nb


def test_whole_pipeline_with_batched_su_equals_su_solve():
    """ADMM loop of the CPU port (float32 state, 8 iterations) with its su-QPs routed through the batched pipeline."""
    from oracle import cpu_port
    import bench
    inp = bench.build_inputs(6, 9000)
    args = (rectangle_robot(), 30, 20, 4, inp['nom_s'], inp['nom_u'], inp['ref_s'], inp['ref_speed'], inp['obs_A'],
            inp['obs_b'], inp['obs_kind'], inp['obs_count'])
    a = cpu_port.solve_batch(*args, iter_num=8, iter_threshold=0.0, threads=2)
    os.environ['RDA_PORT_SU_BATCHED'] = '1'
    try:
        b = cpu_port.solve_batch(*args, iter_num=8, iter_threshold=0.0, threads=2)
    finally:
        del os.environ['RDA_PORT_SU_BATCHED']
    np.testing.assert_allclose(b['u'], a['u'], atol=1e-5)
    np.testing.assert_allclose(b['s'], a['s'], atol=1e-5)
