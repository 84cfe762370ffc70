#!/bin/bash
# final evidence of the round: tests, smoke, bench (own arm + reference arm), launch list, full ncu capture, sanitizer
export RDA_B200_NO_BUILD=1
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.txt 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu_$TAG.txt
tail -3 gpurun_out/pytest_gpu_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.txt 2>&1; tail -1 gpurun_out/smoke_$TAG.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
l=json.load(open('gpurun_out/bench_$TAG.json')); r=json.load(open('gpurun_out/bench_reference_$TAG.json'))
print('value', round(l['value']), 'e2e', round(l['e2e']['value']), l['roofline']['kernel_ms'], l['status_bits'], 'cpu', round(l['cpu_baseline']['value']), l['cpu_baseline']['cores'], 'ref arm', round(r['value']))
print({k: l.get(k) for k in ('single_instance','path_track_control_step','config_B','config_C','early_stop','closed_loop','harsh_geometry','disc_robot')})
PY
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 -c 620 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-probes > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_su|k_cells_mid|k_cells_slow|k_cells_coh|k_cells_fast|k_cells_extra" --launch-skip 30 -c 6 -o gpurun_out/prof_$TAG -f python tools/profile_target.py > gpurun_out/prof_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
for mode in 0 1; do RDA_B200_SMALL=$mode compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/sanitizer_memcheck_small${mode}_$TAG.log 2>&1; tail -2 gpurun_out/sanitizer_memcheck_small${mode}_$TAG.log; done
RDA_B200_SMALL=1 compute-sanitizer --tool racecheck python tools/sanitize_target.py > gpurun_out/sanitizer_racecheck_small1_$TAG.log 2>&1; tail -2 gpurun_out/sanitizer_racecheck_small1_$TAG.log
