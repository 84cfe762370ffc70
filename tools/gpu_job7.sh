#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest7.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest7.txt
tail -4 gpurun_out/r02_pytest7.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err; echo $name; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/r02_$name.json'))
    print(round(l['value']), l.get('roofline',{}).get('kernel_ms'), l.get('counters',{}).get('su_ipm_iterations'), l['status_bits'], l['gpu_launches'], l.get('fp32_vs_fp64_su'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r02_$name.err').read()[-800:])
PY
}
EXTRA=""
run j7_base X=1
EXTRA="--batch 1024"
run j7_b1024 X=1
B="python bench.py --steps 2 --warmup 3"
EXTRA="--config B"; run j7_cfgB X=1
EXTRA="--config C"; run j7_cfgC X=1
EXTRA="--config D --global-batch 512"; run j7_cfgD512 X=1
EXTRA="--config E --global-batch 1024"; run j7_cfgE1024 X=1
