#!/bin/bash
# third session of round 2 (3 GPU-minutes left): throughput of the disc body with the searched closed forms (k_cells_dr_mid),
# per-phase times of one steady-state ADMM iteration, counters
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
timeout 120 python - <<'PY' > gpurun_out/disc_robot_r02.json 2> gpurun_out/disc_robot_r02.err
import json, torch, bench
from rda_planner_b200.rda_solver import RDA_solver
from rda_planner_b200.scenarios import disc_robot
from rda_planner_b200 import _cabi
dev = torch.device('cuda:0')
host = bench.build_inputs(2048, 9000)
rows = []
for Bd in (256, 2048):
    dd = {k: torch.from_numpy(v[:Bd]).to(dev) for k, v in host.items()}
    sv = RDA_solver(bench.T, disc_robot(radius=1.2, wheelbase=2.0, dynamics='diff'), max_edge_num=bench.E, max_obs_num=bench.N, iter_num=bench.ITERS,
                    iter_threshold=0.0, time_print=False, batch=Bd, device=dev)
    def step():
        sv.cold_start()
        return sv.iterative_solve_batch(dd['nom_s'], dd['nom_u'], dd['ref_s'], dd['ref_speed'], dd['obs_A'], dd['obs_b'], dd['obs_kind'], dd['obs_count'], False)
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = step(); r = step(); e1.record(); torch.cuda.synchronize()
    c = sv.state_buffer(_cabi.BUF_COUNTERS).cpu().tolist()
    rows.append({'batch': Bd, 'solves_per_s': 2 * Bd / (e0.elapsed_time(e1) * 1e-3), 'kept_previous_iterate': int((r['status'] & 6).ne(0).sum()),
                 'cells_closed_form': c[0], 'cells_barrier': c[1], 'cells_failed': c[2], 'launches': sv.launch_count()})
print(json.dumps({'what': 'metric shape (T=30, N=20 boxes, 50 iterations, cold start), disc body of radius 1.2 m (cone_type norm2), diff drive', 'rows': rows}))
PY
cat gpurun_out/disc_robot_r02.json; tail -c 400 gpurun_out/disc_robot_r02.err
