#!/bin/bash
# round 2, session 2, job 1: GPU tests of the new rows (disc body, half-space obstacles, cooperative last pass) + the whole GPU suite,
# then knob measurements: cooperative last pass (metric row, small batches, configs C/D/E), su-QP step tolerance variant
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest16.txt 2>&1; echo "pytest rc $?" >> gpurun_out/s2_pytest16.txt
tail -6 gpurun_out/s2_pytest16.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/s2_$name.json 2> gpurun_out/s2_$name.err; echo $name; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/s2_$name.json'))
    print(round(l['value']), l.get('roofline',{}).get('kernel_ms'), l.get('counters'), l.get('status_bits'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/s2_$name.err').read()[-800:])
PY
}
EXTRA=""
run j16_base X=1
run j16_coop RDA_B200_SLOW_COOP=1
run j16_step4 RDA_B200_LIB=$PWD/rda_planner_b200/librda_b200_step4.so
EXTRA="--batch 4096"; run j16_b4096 X=1; run j16_b4096_coop RDA_B200_SLOW_COOP=1
EXTRA="--batch 1024"; run j16_b1024 X=1; run j16_b1024_coop RDA_B200_SLOW_COOP=1
cfg() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/s2_$name.json 2> gpurun_out/s2_$name.err; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/s2_$name.json')); print('$name', round(l['value']), l['ms_per_step'])
except Exception as e:
    print('$name ERR', e); print(open('gpurun_out/s2_$name.err').read()[-600:])
PY
}
EXTRA="--config C"; cfg j16_C X=1; cfg j16_C_coop RDA_B200_SLOW_COOP=1
EXTRA="--config D --global-batch 512"; cfg j16_D512 X=1; cfg j16_D512_coop RDA_B200_SLOW_COOP=1
EXTRA="--config E --global-batch 1024"; cfg j16_E1024 X=1; cfg j16_E1024_coop RDA_B200_SLOW_COOP=1
# disc body probe (metric shape, 2048 instances)
python - <<'PY' > gpurun_out/s2_j16_disc.txt 2>&1
import json, torch, bench
from rda_planner_b200.rda_solver import RDA_solver
from rda_planner_b200.scenarios import disc_robot
dev = torch.device('cuda:0')
for Bd in (256, 2048):
    host = bench.build_inputs(Bd, 9000)
    dd = {k: torch.from_numpy(v).to(dev) for k, v in host.items()}
    sv = RDA_solver(bench.T, disc_robot(radius=1.2, wheelbase=2.0, dynamics='diff'), max_edge_num=bench.E, max_obs_num=bench.N, iter_num=bench.ITERS,
                    iter_threshold=0.0, time_print=False, batch=Bd, device=dev)
    def step():
        sv.cold_start()
        return sv.iterative_solve_batch(dd['nom_s'], dd['nom_u'], dd['ref_s'], dd['ref_speed'], dd['obs_A'], dd['obs_b'], dd['obs_kind'], dd['obs_count'], False)
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = step(); r = step(); e1.record(); torch.cuda.synchronize()
    from rda_planner_b200 import _cabi
    print(json.dumps({'batch': Bd, 'solves_per_s': 2 * Bd / (e0.elapsed_time(e1) * 1e-3), 'status6': int((r['status'] & 6).ne(0).sum()),
                      'counters': sv.state_buffer(_cabi.BUF_COUNTERS).cpu().tolist()}))
PY
cat gpurun_out/s2_j16_disc.txt | tail -3
