#!/bin/bash
# BASELINE configs D (4 GPUs) and E (8 GPUs) at their stated global batch, with and without the rank-0 scatter in the timed region
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 4 --master-port 29611 bench.py --config D --steps 3 --warmup 3 > gpurun_out/r02_cfgD_4gpu.json 2> gpurun_out/r02_cfgD_4gpu.err
$TR --nproc-per-node 4 --master-port 29612 bench.py --config D --steps 3 --warmup 3 --scatter-from-rank0 > gpurun_out/r02_cfgD_4gpu_scatter.json 2> gpurun_out/r02_cfgD_4gpu_scatter.err
$TR --nproc-per-node 8 --master-port 29613 bench.py --config E --steps 2 --warmup 3 > gpurun_out/r02_cfgE_8gpu.json 2> gpurun_out/r02_cfgE_8gpu.err
$TR --nproc-per-node 8 --master-port 29614 bench.py --config E --steps 2 --warmup 3 --scatter-from-rank0 > gpurun_out/r02_cfgE_8gpu_scatter.json 2> gpurun_out/r02_cfgE_8gpu_scatter.err
python bench.py --config D --steps 3 --warmup 3 > gpurun_out/r02_cfgD_1gpu.json 2> gpurun_out/r02_cfgD_1gpu.err
for f in cfgD_4gpu cfgD_4gpu_scatter cfgE_8gpu cfgE_8gpu_scatter cfgD_1gpu; do echo $f; python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/r02_$f.json').read().splitlines() if x.startswith('{')][-1])
    print(round(l['value']), 'ms/step', round(l['ms_per_step'],1), l['n_gpus'], l['config']['batch_per_gpu'], l['status_bits'], l.get('scatter_bytes_per_step'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r02_$f.err').read()[-600:])
PY
done
