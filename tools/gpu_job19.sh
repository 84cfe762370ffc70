#!/bin/bash
# round 2, session 2, job 4: last pass split into thread-per-cell closed forms (k_cells_extra) + cooperative interior point
# pass on the remainder; cooperative two-cone barrier for the disc body
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s2_pytest19.txt 2>&1; echo "pytest rc $?" >> gpurun_out/s2_pytest19.txt
tail -6 gpurun_out/s2_pytest19.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/s2_$name.json 2> gpurun_out/s2_$name.err; echo $name; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/s2_$name.json'))
    print(round(l['value'],1), l.get('roofline',{}).get('kernel_ms'), l.get('counters'), l.get('status_bits'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/s2_$name.err').read()[-800:])
PY
}
EXTRA=""
run j19_base X=1
run j19_coop0 RDA_B200_SLOW_COOP=0
run j19_ctas8 RDA_B200_SLOW_CTAS=8
EXTRA="--batch 1024"; run j19_b1024 X=1
cfg() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/s2_$name.json 2> gpurun_out/s2_$name.err; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/s2_$name.json')); print('$name', round(l['value']), l['ms_per_step'])
except Exception as e:
    print('$name ERR', e); print(open('gpurun_out/s2_$name.err').read()[-600:])
PY
}
EXTRA="--config C"; cfg j19_C X=1
EXTRA="--config D --global-batch 512"; cfg j19_D512 X=1
python - <<'PY' > gpurun_out/s2_j19_disc.txt 2>&1
import json, os, torch, bench
from rda_planner_b200.rda_solver import RDA_solver
from rda_planner_b200.scenarios import disc_robot
from rda_planner_b200 import _cabi
dev = torch.device('cuda:0')
host = bench.build_inputs(2048, 9000)
for coop in ('1', '0'):
    os.environ['RDA_B200_DR_COOP'] = coop
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
        print(json.dumps({'dr_coop': coop, 'batch': Bd, 'solves_per_s': 2 * Bd / (e0.elapsed_time(e1) * 1e-3), 'status6': int((r['status'] & 6).ne(0).sum()),
                          'counters': sv.state_buffer(_cabi.BUF_COUNTERS).cpu().tolist()}), flush=True)
PY
cat gpurun_out/s2_j19_disc.txt | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 56 --csv --log-file gpurun_out/s2_launches_19.csv env RDA_B200_SPLIT_MIN=1000000 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-probes > gpurun_out/s2_ncu_19.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/s2_launches_19.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    k = r[ki].split('(')[0][-40:]
    agg[k][0] += 1; agg[k][1] += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]): print(f'{k:42s} n={n:4d} total={t/1e3:9.1f} us  avg={t/n/1e3:8.1f} us')
PY
