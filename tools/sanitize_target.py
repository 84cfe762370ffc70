"""Small target for compute-sanitizer (memcheck / racecheck): a short batched solve covering polygon and
disc obstacles, static and moving."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from rda_planner_b200.rda_solver import RDA_solver, pack_obstacles  # noqa: E402
from rda_planner_b200.scenarios import rectangle_robot, make_instance  # noqa: E402

for kind, moving in (('polygon', False), ('circle', True)):
    T, N, B = 10, 5, 48
    car = rectangle_robot()
    insts = [make_instance(40 + i, T=T, N=N, E=4, kind=kind, moving=moving, lateral=(0.3, 3.5)) for i in range(B)]
    packs = [pack_obstacles(list(i['obstacles']), T, N, 4) for i in insts]
    tv = any(p[4] for p in packs)
    if tv:      # all instances must share the layout: expand static ones
        packs = [(np.repeat(p[0], T + 1, 1) if not p[4] else p[0], np.repeat(p[1], T + 1, 1) if not p[4] else p[1], p[2], p[3], True)
                 for p in packs]
    g = RDA_solver(T, car, 4, N, iter_num=4, iter_threshold=0.0, time_print=False, batch=B)
    out = g.iterative_solve_batch(np.stack([i['nom_s'] for i in insts]), np.stack([i['nom_u'] for i in insts]),
                                  np.stack([i['ref'] for i in insts]), np.array([4.0] * B), np.stack([p[0] for p in packs]),
                                  np.stack([p[1] for p in packs]), np.stack([p[2] for p in packs]),
                                  np.array([p[3] for p in packs]), tv)
    torch.cuda.synchronize()
    print(kind, 'ok', bool(torch.isfinite(out['u']).all()), int((out['status'] & 6).sum()))
