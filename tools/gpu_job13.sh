#!/bin/bash
export RDA_B200_NO_BUILD=1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
for c in 0 11 10 9; do RDA_B200_SU_MAXCTAS=$c $B > gpurun_out/r02_j13_$c.json 2>gpurun_out/r02_j13_$c.err; python -c "
import json; l=json.load(open('gpurun_out/r02_j13_$c.json')); print('maxctas $c', round(l['value']), l['ms_per_step'], l['roofline']['kernel_ms'])" || tail -3 gpurun_out/r02_j13_$c.err; done
for c in 11 10; do RDA_B200_SPLIT_PARTS=4 RDA_B200_SU_MAXCTAS=$c $B > gpurun_out/r02_j13_p4_$c.json 2>/dev/null; python -c "
import json; l=json.load(open('gpurun_out/r02_j13_p4_$c.json')); print('parts4 maxctas $c', round(l['value']), l['ms_per_step'])"; done
