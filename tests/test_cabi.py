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
    assert ctypes.sizeof(_cabi.Config) == 4 * (8 + 2 + 2 + 2 + 2 + 16 + 8)
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
