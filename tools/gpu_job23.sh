#!/bin/bash
# verification of the row preload in the searched / rare closed-form passes (+ edge-contact closed forms of the disc body): GPU suite, bench
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02.txt 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_r02.txt
tail -4 gpurun_out/pytest_gpu_r02.txt
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
python - <<PY
import json
l=json.load(open('gpurun_out/bench_r02.json'))
print('value', round(l['value']), 'e2e', round(l['e2e']['value']), l['roofline']['kernel_ms'], l['status_bits'], 'cpu', round(l['cpu_baseline']['value']), l['cpu_baseline']['cores'])
print({k: l.get(k) for k in ('single_instance','path_track_control_step','config_B','config_C','early_stop','harsh_geometry','disc_robot')})
PY
