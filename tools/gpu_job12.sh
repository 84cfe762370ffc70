#!/bin/bash
export RDA_B200_NO_BUILD=1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
for v in ch3 ch4; do RDA_B200_LIB=$PWD/rda_planner_b200/librda_b200_$v.so $B > gpurun_out/r02_j12_$v.json 2> gpurun_out/r02_j12_$v.err; echo $v; python -c "
import json; l=json.load(open('gpurun_out/r02_j12_$v.json')); print(round(l['value']), l['roofline']['kernel_ms'])" || tail -5 gpurun_out/r02_j12_$v.err; done
$B > gpurun_out/r02_j12_ch2.json 2>/dev/null; python -c "
import json; l=json.load(open('gpurun_out/r02_j12_ch2.json')); print('ch2', round(l['value']), l['roofline']['kernel_ms'])"
