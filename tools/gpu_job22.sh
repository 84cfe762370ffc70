#!/bin/bash
# final single-GPU numbers of the BASELINE configs B-E (strong-scaling side benches)
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
cfg() { name=$1; shift; python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/$name.json')); print('$name', round(l['value']), round(l['ms_per_step'],1), l.get('status_bits'))
except Exception as e:
    print('$name ERR', e); print(open('gpurun_out/$name.err').read()[-600:])
PY
}
cfg config_B_r02 --config B
cfg config_C_r02 --config C
cfg config_D_1gpu_r02 --config D
cfg config_D_1gpu_b512_r02 --config D --global-batch 512
cfg config_E_1gpu_b1024_r02 --config E --global-batch 1024
