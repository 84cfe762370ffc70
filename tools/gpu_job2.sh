#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest2.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest2.txt
tail -3 gpurun_out/r02_pytest2.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err; echo $name; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/r02_$name.json'))
    print(round(l['value']), l['roofline']['kernel_ms'], l['counters'], l['status_bits'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r02_$name.err').read()[-600:])
PY
}
EXTRA=""
run base2 X=1
run lean2b RDA_B200_LEAN2=1
EXTRA="--batch 1024"
run b1024b X=1
run b1024b_lean2 RDA_B200_LEAN2=1
