#!/bin/bash
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest10.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest10.txt
tail -4 gpurun_out/r02_pytest10.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke10.txt 2>&1; tail -2 gpurun_out/r02_smoke10.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err; echo $name; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/r02_$name.json'))
    print(round(l['value']), l.get('roofline',{}).get('kernel_ms'), l.get('counters'), l['status_bits'], {k: l.get(k) for k in ('single_instance','path_track_control_step','config_B','config_C','early_stop')})
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r02_$name.err').read()[-800:])
PY
}
EXTRA=""
run j10_base X=1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run j10_prune05 RDA_B200_SU_PRUNE=0.5
EXTRA="--batch 64";  run j10_b64_small RDA_B200_SMALL=1;  run j10_b64_stream RDA_B200_SMALL=0
EXTRA="--batch 296"; run j10_b296_small RDA_B200_SMALL=1; run j10_b296_stream RDA_B200_SMALL=0
EXTRA="--batch 1024"; run j10_b1024_small RDA_B200_SMALL=1; run j10_b1024_stream RDA_B200_SMALL=0
