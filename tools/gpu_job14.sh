#!/bin/bash
export RDA_B200_NO_BUILD=1
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_pytest14.txt; cat gpurun_out/r02_pytest14.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run() { name=$1; shift; env "$@" $B > gpurun_out/r02_j14_$name.json 2>gpurun_out/r02_j14_$name.err; python -c "
import json; l=json.load(open('gpurun_out/r02_j14_$name.json')); print('$name', round(l['value']), l['ms_per_step'], l['roofline']['kernel_ms'], l.get('counters'), l.get('status_bits'))" || tail -3 gpurun_out/r02_j14_$name.err; }
run base X=1
run cpw8 RDA_B200_SLOW_CPW=8
run cpw2 RDA_B200_SLOW_CPW=2
run ctas8 RDA_B200_SLOW_CTAS=8
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes --config C > gpurun_out/r02_j14_C.json 2>gpurun_out/r02_j14_C.err; python -c "
import json; l=json.load(open('gpurun_out/r02_j14_C.json')); print('C', round(l['value']), l['ms_per_step'])"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes --batch 1024 > gpurun_out/r02_j14_b1024.json 2>gpurun_out/r02_j14_b1024.err; python -c "
import json; l=json.load(open('gpurun_out/r02_j14_b1024.json')); print('b1024', round(l['value']), l['ms_per_step'], l['roofline']['kernel_ms'])"
