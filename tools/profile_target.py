"""Short target for `ncu --set full`: one batched solve of a few ADMM iterations at the bench batch size.
The batch is kept whole (no two-stream split) so that each captured launch is the full-batch kernel."""
import json
import os
import sys

os.environ.setdefault('RDA_B200_SPLIT_MIN', '1000000000')

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ['bench']
import bench  # noqa: E402
import torch  # noqa: E402
from rda_planner_b200.rda_solver import RDA_solver  # noqa: E402
from rda_planner_b200.scenarios import rectangle_robot  # noqa: E402

B = int(os.environ.get('PROF_BATCH', '16384'))
ITERS = 12
host = bench.build_inputs(B, 9000)
dev = torch.device('cuda:0')
d = {k: torch.from_numpy(v).to(dev) for k, v in host.items()}
s = RDA_solver(30, rectangle_robot(), max_edge_num=4, max_obs_num=20, iter_num=ITERS, iter_threshold=0.0, time_print=False, batch=B)
s.iterative_solve_batch(d['nom_s'], d['nom_u'], d['ref_s'], d['ref_speed'], d['obs_A'], d['obs_b'], d['obs_kind'], d['obs_count'], False)
torch.cuda.synchronize()
os.makedirs('gpurun_out', exist_ok=True)
json.dump({'batch': B, 'admm_iterations': ITERS, 'captured': 'one launch of each kernel of an ADMM iteration in steady state'},
          open('gpurun_out/prof_r01_meta.json', 'w'))
