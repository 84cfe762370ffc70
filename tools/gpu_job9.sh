#!/bin/bash
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest9.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest9.txt
tail -4 gpurun_out/r02_pytest9.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err; echo $name; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/r02_$name.json'))
    print(round(l['value']), l.get('roofline',{}).get('kernel_ms'), l.get('counters'), l['status_bits'], l.get('counters_rank0_last_step'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r02_$name.err').read()[-800:])
PY
}
EXTRA=""
run j9_base X=1
run j9_noprune RDA_B200_SU_PRUNE=0
run j9_prune05 RDA_B200_SU_PRUNE=0.5
run j9_prune2 RDA_B200_SU_PRUNE=2.0
EXTRA="--batch 1024"
run j9_b1024 X=1
B="python bench.py --steps 1 --warmup 3"
EXTRA="--config C"; run j9_cfgC X=1
ncu --set full --clock-control none --import-source on -k regex:"k_su" --launch-skip 6 -c 1 -o gpurun_out/prof_r02b -f python tools/profile_target.py > gpurun_out/prof_r02b.log 2>&1
ls -la gpurun_out/prof_r02b.ncu-rep
