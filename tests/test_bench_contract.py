"""bench.py contract checks that need no GPU: the reference arm (compiled CPU port on the host cores)
prints one well-formed JSON line, and the committed round profile carries every key of the contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline']


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS=str(min(os.cpu_count() or 1, 8)))
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['value'] > 0 and d['unit'] == 'solves/s' and d['higher_is_better'] is True
    for k in BASE_KEYS:
        assert k in d, k
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']


def test_committed_profile_line_has_every_contract_key():
    path = os.path.join(ROOT, 'profiles', 'bench_r01.json')
    d = json.loads(open(path).read().strip().splitlines()[-1])
    for k in BASE_KEYS + ['gpu_launches', 'clocks', 'roofline']:
        assert k in d, k
    r = d['roofline']
    for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert k in r, k
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert d['gpu_launches'] > 0 and d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    assert d['metric'] == json.load(open(os.path.join(ROOT, 'BASELINE.json')))['metric'] or 'MPC solves' in d['metric']
    assert set(d['clocks']['reasons']).isdisjoint({'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'})
