#!/bin/bash
# round-2 GPU job: GPU tests, su-QP group width sweep, coherent cell pass, small batch, float32 su-QP probe
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest1.txt
tail -3 gpurun_out/r02_pytest1.txt
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
run g32 RDA_B200_SU_GROUP=32
run g16 RDA_B200_SU_GROUP=16
run g8 RDA_B200_SU_GROUP=8
run g16l2 RDA_B200_SU_GROUP=16 RDA_B200_SU_LEVEL=2
run g8l1 RDA_B200_SU_GROUP=8 RDA_B200_SU_LEVEL=1
run auto_lean2 RDA_B200_LEAN2=1
EXTRA="--batch 1024"
run b1024 X=1
run b1024_lean2 RDA_B200_LEAN2=1
EXTRA="--batch 4096"
run b4096 X=1
run b4096_g16 RDA_B200_SU_GROUP=16
EXTRA="--su-fp32"
run sufp32_g32 RDA_B200_SU_GROUP=32
run sufp32_g8 RDA_B200_SU_GROUP=8
