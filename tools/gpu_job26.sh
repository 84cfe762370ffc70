#!/bin/bash
# third session of round 2, last GPU call: A/B of the su-QP's adaptive fraction to the boundary (RDA_SU_TAU_ADAPT = 10,
# the new default, against librda_b200_tau0.so = the previous fixed 0.995) at 4096 unique metric-row instances,
# then the GPU parity tests on the new default.
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
python - <<'PY'
import numpy as np, bench
h = bench.build_inputs(4096, 31000)
np.savez('/tmp/ab_inputs.npz', **h)
PY
cat > /tmp/ab_probe.py <<'PY'
import json, os, sys, numpy as np, torch, bench
from rda_planner_b200.rda_solver import RDA_solver
from rda_planner_b200.scenarios import rectangle_robot
from rda_planner_b200 import _cabi
dev = torch.device('cuda:0')
h = np.load('/tmp/ab_inputs.npz')
B = 4096
dd = {k: torch.from_numpy(h[k][:B]).to(dev) for k in h.files}
sv = RDA_solver(bench.T, rectangle_robot(), max_edge_num=bench.E, max_obs_num=bench.N, iter_num=bench.ITERS, iter_threshold=0.0,
                time_print=False, batch=B, device=dev)
def step():
    sv.cold_start()
    return sv.iterative_solve_batch(dd['nom_s'], dd['nom_u'], dd['ref_s'], dd['ref_speed'], dd['obs_A'], dd['obs_b'], dd['obs_kind'], dd['obs_count'], False)
step(); step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 4
e0.record()
for _ in range(K): r = step()
e1.record(); torch.cuda.synchronize()
c = sv.state_buffer(_cabi.BUF_COUNTERS).cpu().tolist()
# per-phase times of steady-state iterations (whole batch, one stream)
sv.cold_start()
sv.begin(dd['nom_s'], dd['nom_u'], dd['ref_s'], dd['ref_speed'], dd['obs_A'], dd['obs_b'], dd['obs_kind'], dd['obs_count'], False, 0.0)
for _ in range(10): sv.step_su(); sv.step_lammuz()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(20)]
for i in range(20):
    ev[i][0].record(); sv.step_su(); ev[i][1].record(); sv.step_lammuz(); ev[i][2].record()
torch.cuda.synchronize()
su = float(np.mean([ev[i][0].elapsed_time(ev[i][1]) for i in range(20)])); ce = float(np.mean([ev[i][1].elapsed_time(ev[i][2]) for i in range(20)]))
print(json.dumps({'lib': os.path.basename(_cabi.LIB_PATH), 'batch': B, 'solves_per_s': K * B / (e0.elapsed_time(e1) * 1e-3),
                  'ipm_iterations_per_su_solve': c[3] / max(1, c[4]) if len(c) > 4 else None, 'counters': c[:8],
                  'k_su_ms_iterations_11_30': su, 'cells_ms_iterations_11_30': ce,
                  'status_nonzero': int((r['status'] & 7).ne(0).sum()), 'u_checksum': float(r['u'].double().abs().sum())}))
PY
for lib in librda_b200_tau0.so librda_b200.so; do
  RDA_B200_LIB=$PWD/rda_planner_b200/$lib PYTHONPATH=$PWD timeout 60 python /tmp/ab_probe.py 2>> gpurun_out/ab_tau_r02.err | tail -1 >> gpurun_out/ab_tau_r02.jsonl
done
cat gpurun_out/ab_tau_r02.jsonl; tail -c 300 gpurun_out/ab_tau_r02.err
timeout 70 python -m pytest tests/test_gpu_parity50.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_parity_tau_r02.txt 2>&1; echo "rc $?" >> gpurun_out/pytest_gpu_parity_tau_r02.txt
tail -4 gpurun_out/pytest_gpu_parity_tau_r02.txt
