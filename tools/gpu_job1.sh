#!/bin/bash
# round-2 first GPU job: GPU tests of the merged tree, baseline vs coherent cell pass, small batch, float32 su-QP probe
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest1.txt
tail -3 gpurun_out/r02_pytest1.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
$B > gpurun_out/r02_base.json 2> gpurun_out/r02_base.err
RDA_B200_LEAN2=1 $B > gpurun_out/r02_lean2.json 2> gpurun_out/r02_lean2.err
$B --batch 1024 > gpurun_out/r02_base_b1024.json 2> gpurun_out/r02_base_b1024.err
RDA_B200_LEAN2=1 $B --batch 1024 > gpurun_out/r02_lean2_b1024.json 2> gpurun_out/r02_lean2_b1024.err
$B --su-fp32 > gpurun_out/r02_sufp32.json 2> gpurun_out/r02_sufp32.err
for f in base lean2 base_b1024 lean2_b1024 sufp32; do echo $f; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/r02_$f.json'))
    print(round(l['value']), l['roofline']['kernel_ms'], l['counters'], l['status_bits'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r02_$f.err').read()[-800:])
PY
done
