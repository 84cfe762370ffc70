#!/bin/bash
# third session of round 2, last GPU call: occupancy variants of k_su (box slacks in global memory + register cap for 14 / 16
# resident one-warp CTAs per SM instead of 12) against the default build, 9472 = 148 x 64 unique metric-row instances;
# every variant's results compared BITWISE with the default build's.
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
python - <<'PY'
import numpy as np, bench
h = bench.build_inputs(9472, 31000)
np.savez('/tmp/ab_inputs.npz', **h)
PY
cat > /tmp/ab_probe.py <<'PY'
import json, os, sys, numpy as np, torch, bench
from rda_planner_b200.rda_solver import RDA_solver
from rda_planner_b200.scenarios import rectangle_robot
from rda_planner_b200 import _cabi
dev = torch.device('cuda:0')
h = np.load('/tmp/ab_inputs.npz')
B = 9472
dd = {k: torch.from_numpy(h[k][:B]).to(dev) for k in h.files}
sv = RDA_solver(bench.T, rectangle_robot(), max_edge_num=bench.E, max_obs_num=bench.N, iter_num=bench.ITERS, iter_threshold=0.0,
                time_print=False, batch=B, device=dev)
def step():
    sv.cold_start()
    return sv.iterative_solve_batch(dd['nom_s'], dd['nom_u'], dd['ref_s'], dd['ref_speed'], dd['obs_A'], dd['obs_b'], dd['obs_kind'], dd['obs_count'], False)
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 3
e0.record()
for _ in range(K): r = step()
e1.record(); torch.cuda.synchronize()
c = sv.state_buffer(_cabi.BUF_COUNTERS).cpu().tolist()
res = {k: r[k].cpu().numpy() for k in ('u', 's', 'status')}
tag = os.path.basename(_cabi.LIB_PATH)
same = None
if os.path.exists('/tmp/ab_base.npz'):
    z = np.load('/tmp/ab_base.npz')
    same = {k: bool(np.array_equal(z[k], res[k])) for k in res}
    same['max_abs_u_diff'] = float(np.abs(z['u'] - res['u']).max())
else:
    np.savez('/tmp/ab_base.npz', **res)
sv.cold_start()
sv.begin(dd['nom_s'], dd['nom_u'], dd['ref_s'], dd['ref_speed'], dd['obs_A'], dd['obs_b'], dd['obs_kind'], dd['obs_count'], False, 0.0)
for _ in range(10): sv.step_su(); sv.step_lammuz()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(20)]
for i in range(20):
    ev[i][0].record(); sv.step_su(); ev[i][1].record(); sv.step_lammuz(); ev[i][2].record()
torch.cuda.synchronize()
su = float(np.mean([ev[i][0].elapsed_time(ev[i][1]) for i in range(20)])); ce = float(np.mean([ev[i][1].elapsed_time(ev[i][2]) for i in range(20)]))
print(json.dumps({'lib': tag, 'batch': B, 'solves_per_s': K * B / (e0.elapsed_time(e1) * 1e-3), 'k_su_ms_whole_batch_one_stream': su,
                  'cells_ms_whole_batch_one_stream': ce, 'ipm_iterations_per_su_solve': c[3] / max(1, c[4]),
                  'status_nonzero': int((r['status'] & 7).ne(0).sum()), 'bitwise_equal_to_default_build': same}))
PY
for lib in librda_b200.so variants/lib_bsg16.so variants/lib_bsg14.so variants/lib_r16.so variants/lib_bsg16ch3.so; do
  RDA_B200_LIB=$PWD/rda_planner_b200/$lib PYTHONPATH=$PWD timeout 40 python /tmp/ab_probe.py 2>> gpurun_out/ab_occupancy_r02.err | tail -1 >> gpurun_out/ab_occupancy_r02.jsonl
done
cat gpurun_out/ab_occupancy_r02.jsonl; tail -c 300 gpurun_out/ab_occupancy_r02.err
RDA_B200_LIB=$PWD/rda_planner_b200/variants/lib_bsg16.so timeout 45 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_parity_bsg16_r02.txt 2>&1; echo "rc $?" >> gpurun_out/pytest_gpu_parity_bsg16_r02.txt
tail -3 gpurun_out/pytest_gpu_parity_bsg16_r02.txt
