#!/bin/bash
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "graph or batched or trajectory" > gpurun_out/r02_pytest8.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest8.txt
tail -3 gpurun_out/r02_pytest8.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err; echo $name; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/r02_$name.json'))
    print(round(l['value']), l.get('roofline',{}).get('kernel_ms'), l['status_bits'], l.get('counters_rank0_last_step'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r02_$name.err').read()[-800:])
PY
}
EXTRA=""
run j8_cpw8 RDA_B200_SLOW_CPW=8 RDA_B200_SLOW_CTAS=32
run j8_cpw2 RDA_B200_SLOW_CPW=2 RDA_B200_SLOW_CTAS=64
run j8_cpw1 RDA_B200_SLOW_CPW=1 RDA_B200_SLOW_CTAS=128
EXTRA="--batch 1024"
run j8_b1024_cpw2 RDA_B200_SLOW_CPW=2 RDA_B200_SLOW_CTAS=64
run j8_b1024_cpw1 RDA_B200_SLOW_CPW=1 RDA_B200_SLOW_CTAS=64
B="python bench.py --steps 1 --warmup 3"
EXTRA="--config C --global-batch 128"; run j8_cfgC X=1
EXTRA="--config C --global-batch 128"; run j8_cfgC_cpw1 RDA_B200_SLOW_CPW=1 RDA_B200_SLOW_CTAS=128
