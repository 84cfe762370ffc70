#!/bin/bash
# round 2, session 2, job 2: whole GPU suite (cooperative last pass now the default), k_cells_mid grid / occupancy knobs,
# cooperative interior point pass inside the persistent small-batch kernel
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s2_pytest17.txt 2>&1; echo "pytest rc $?" >> gpurun_out/s2_pytest17.txt
tail -6 gpurun_out/s2_pytest17.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/s2_$name.json 2> gpurun_out/s2_$name.err; echo $name; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/s2_$name.json'))
    print(round(l['value'],1), l.get('roofline',{}).get('kernel_ms'), l.get('status_bits'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/s2_$name.err').read()[-800:])
PY
}
EXTRA=""
run j17_base X=1
run j17_mid6 RDA_B200_MID_CTAS=6
run j17_mid12 RDA_B200_MID_CTAS=12
run j17_mid18 RDA_B200_MID_CTAS=18
run j17_mid8lib RDA_B200_LIB=$PWD/rda_planner_b200/librda_b200_mid8.so
run j17_mid8lib16 RDA_B200_LIB=$PWD/rda_planner_b200/librda_b200_mid8.so RDA_B200_MID_CTAS=16
EXTRA="--batch 1024"; run j17_b1024 X=1; run j17_b1024_mid6 RDA_B200_MID_CTAS=6
for b in 1 64 296; do EXTRA="--batch $b"; run j17_b${b}_smallcoop1 X=1; run j17_b${b}_smallcoop0 RDA_B200_SMALL_COOP=0; done
