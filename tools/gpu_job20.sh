#!/bin/bash
# round 2, session 2: final bench line again after the host-side fix of the single-instance API (obstacle packing cache)
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
python - <<PY
import json
l=json.load(open('gpurun_out/bench_r02.json'))
print('value', round(l['value']), 'e2e', round(l['e2e']['value']), l['roofline']['kernel_ms'], l['status_bits'], 'cpu', round(l['cpu_baseline']['value']), l['cpu_baseline']['cores'])
print({k: l.get(k) for k in ('single_instance','path_track_control_step','config_B','config_C','early_stop','harsh_geometry','disc_robot')})
PY
