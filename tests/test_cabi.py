"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol that
include/rda_b200.h declares.  No compute call is made here."""
import ctypes
import os
import re

from rda_planner_b200 import build as rbuild
from rda_planner_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    so = rbuild.build()
    lib = ctypes.CDLL(so)
    header = open(os.path.join(ROOT, 'include', 'rda_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    names = set(re.findall(r'\b(rda_[a-z_]+)\s*\(', header))
    assert len(names) >= 14
    for n in sorted(names):
        assert hasattr(lib, n), f'{n} declared in include/rda_b200.h but not exported'
    assert set(_cabi.EXPORTS) == names
    lib.rda_version.restype = ctypes.c_char_p
    assert b'sm_100a' in lib.rda_version()


def test_struct_layouts_match_header():
    # sizes implied by the header (4-byte fields, no padding)
    assert ctypes.sizeof(_cabi.Config) == 4 * (8 + 2 + 2 + 2 + 2 + 16 + 8 + 1)      # ... + robot_cone
    assert ctypes.sizeof(_cabi.Tunables) == 24
    assert ctypes.sizeof(_cabi.Inputs) == 8 * 8 + 8
    assert ctypes.sizeof(_cabi.Outputs) == 6 * 8


def test_no_cpu_fallback_without_gpu():
    import torch
    import pytest
    from rda_planner_b200.scenarios import rectangle_robot
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from rda_planner_b200.rda_solver import RDA_solver
    with pytest.raises(RuntimeError):
        RDA_solver(5, rectangle_robot(), 4, 2)


def test_product_never_imports_oracle():
    for pkg in ('rda_planner_b200', 'RDA_planner'):
        for f in os.listdir(os.path.join(ROOT, pkg)):
            if f.endswith('.py'):
                src = open(os.path.join(ROOT, pkg, f)).read()
                assert 'oracle' not in src.replace('# oracle', ''), f'{pkg}/{f} mentions oracle'


def test_usage_errors_are_return_codes_not_exceptions():
    """Error behaviour of the boundary (include/rda_b200.h): usage errors come back as negative codes
    before any device work is attempted, so this runs without a GPU."""
    lib = _cabi.load()
    vp = ctypes.c_void_p
    h = vp()
    tun = _cabi.Tunables(8, 1.0, 0.1, 200, 1, 0.5)
    cfg = _cabi.Config()
    assert lib.rda_create(None, ctypes.byref(tun), ctypes.byref(h)) == -1
    cfg.batch, cfg.receding, cfg.max_obs_num, cfg.max_edge_num, cfg.robot_edges = 0, 10, 4, 4, 4
    assert lib.rda_create(ctypes.byref(cfg), ctypes.byref(tun), ctypes.byref(h)) == -1          # batch < 1
    cfg.batch, cfg.max_edge_num = 1, 9
    assert lib.rda_create(ctypes.byref(cfg), ctypes.byref(tun), ctypes.byref(h)) == -2          # > RDA_MAX_EDGE
    cfg.max_edge_num, cfg.dynamics = 4, 7
    assert lib.rda_create(ctypes.byref(cfg), ctypes.byref(tun), ctypes.byref(h)) == -1          # unknown dynamics
    cfg.dynamics = 0                                                                              # G, h all zero:
    assert lib.rda_create(ctypes.byref(cfg), ctypes.byref(tun), ctypes.byref(h)) == -2          # not a polygon
    # disc body (RDA_ROBOT_DISC, cone_type 'norm2'): only G = [[1,0],[0,1],[0,0]], h[2] < 0, three rows; unknown cone refused
    cfg.robot_cone, cfg.robot_edges = _cabi.ROBOT_DISC, 3
    assert lib.rda_create(ctypes.byref(cfg), ctypes.byref(tun), ctypes.byref(h)) == -2          # G all zero
    for k, v in enumerate((1.0, 0.0, 0.0, 1.0, 0.0, 0.0)):
        cfg.G[k] = v
    cfg.h[2] = 0.5
    assert lib.rda_create(ctypes.byref(cfg), ctypes.byref(tun), ctypes.byref(h)) == -2          # radius -h[2] <= 0
    cfg.h[2], cfg.robot_edges = -0.5, 4
    assert lib.rda_create(ctypes.byref(cfg), ctypes.byref(tun), ctypes.byref(h)) == -2          # a disc has three rows
    cfg.robot_edges, cfg.robot_cone = 3, 5
    assert lib.rda_create(ctypes.byref(cfg), ctypes.byref(tun), ctypes.byref(h)) == -2          # unknown cone
    assert h.value is None
    for f in (lib.rda_destroy, ):
        assert f(None) == -1
    assert lib.rda_solve(None, None, None, 1, 0.0, None) == -1
    assert lib.rda_step_su(None, None) == -1 and lib.rda_step_lammuz(None, None) == -1
    assert lib.rda_last_launch_count(None) == -1
    # front end: sizes, missing pointers, more raw shapes than RDA_MAX_SHAPES
    assert lib.rda_pre_process(0, 10, 0, 0.1, 3.0, None, None, None, None, 5, None, 0.1, 10, None, None, None, None) == -1
    assert lib.rda_pre_process(4, 10, 0, 0.1, 3.0, None, None, None, None, 5, None, 0.1, 10, None, None, None, None) == -1
    assert lib.rda_convert_obstacles(4, _cabi.MAX_SHAPES + 1, 5, 10, 4, 0.1, 0, 0, *([None] * 12)) == -2
    assert lib.rda_convert_obstacles(4, 8, 5, 10, 2, 0.1, 0, 0, *([None] * 12)) == -2             # E < 3
    assert lib.rda_convert_obstacles(4, 8, 5, 10, 4, 0.1, 0, 0, *([None] * 12)) == -1
    assert lib.rda_post_process(4, 10, 100, 1, None, None, None, None, None) == -1
    assert lib.rda_motion_predict(4, 10, 5, 0.1, 3.0, None, None, None) == -1
