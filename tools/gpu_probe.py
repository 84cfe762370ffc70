"""Development probe (run under gpurun): GPU solver vs oracle on a few instances; prints gaps."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rda_planner_b200.scenarios import rectangle_robot, make_instance
from rda_planner_b200.rda_solver import RDA_solver
from rda_planner_b200 import _cabi
from oracle.rda_oracle import OracleRDA

def run(seed, T, N, iters, lateral, kind='polygon', dyn='acker', fp64=True):
    car = rectangle_robot(dynamics=dyn)
    inst = make_instance(seed, T=T, N=N, E=4, lateral=lateral, kind=kind, dynamics=dyn)
    ref = [inst['ref'][:, t:t+1] for t in range(T+1)]
    o = OracleRDA(T, car, max_edge_num=4, max_obs_num=N, iter_num=iters, iter_threshold=0.0, cell_solver='geo')
    t0 = time.time()
    uo, io = o.iterative_solve(inst['nom_s'], inst['nom_u'], ref, inst['ref_speed'], list(inst['obstacles']))
    to = time.time() - t0
    g = RDA_solver(T, car, max_edge_num=4, max_obs_num=N, iter_num=iters, iter_threshold=0.0, time_print=False, su_fp64=fp64)
    t0 = time.time()
    ug, ig = g.iterative_solve(inst['nom_s'], inst['nom_u'], ref, inst['ref_speed'], list(inst['obstacles']))
    tg = time.time() - t0
    so = np.hstack(io['opt_state_list']); sg = np.hstack(ig['opt_state_list'])
    cnt = g.state_buffer(_cabi.BUF_COUNTERS).cpu().numpy()
    print(f'seed {seed} T{T} N{N} it{iters} {kind} {dyn} fp64={fp64}: du {np.abs(uo-ug).max():.2e} ds {np.abs(so-sg).max():.2e} '
          f'resi o=({io["resi_dual"]:.3e},{io["resi_pri"]:.3e}) g=({ig["resi_dual"]:.3e},{ig["resi_pri"]:.3e}) status {ig["status"]} '
          f'counters {cnt.tolist()} t_oracle {to:.2f}s t_gpu {tg*1e3:.1f}ms oracle_cells {o.cell_stats}', flush=True)

if __name__ == '__main__':
    print(torch.cuda.get_device_name(0))
    run(3, 10, 4, 1, (1.8, 6.0))
    run(3, 10, 4, 4, (1.8, 6.0))
    run(3, 10, 4, 4, (1.8, 6.0), fp64=False)
    run(11, 10, 5, 1, (0.0, 2.5))
    run(11, 10, 5, 2, (0.0, 2.5))
    run(11, 10, 5, 6, (0.0, 2.5))
    run(12, 10, 5, 6, (0.5, 3.0))
    run(13, 10, 5, 6, (0.5, 3.0), kind='circle')
    run(14, 10, 5, 6, (0.5, 3.0), dyn='diff')
    run(15, 10, 5, 6, (0.5, 3.0), dyn='omni')
    run(16, 20, 10, 8, (0.5, 4.0))
