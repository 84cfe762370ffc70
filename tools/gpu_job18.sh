#!/bin/bash
# round 2, session 2, job 3: GPU suite with the float32-resolution acceptance of edge-contact roots switched off (parity),
# its cost in throughput (variant library with the acceptance on), per-kernel launch list
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s2_pytest18.txt 2>&1; echo "pytest rc $?" >> gpurun_out/s2_pytest18.txt
tail -6 gpurun_out/s2_pytest18.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/s2_$name.json 2> gpurun_out/s2_$name.err; echo $name; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/s2_$name.json'))
    print(round(l['value'],1), l.get('roofline',{}).get('kernel_ms'), l.get('counters'), l.get('status_bits'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/s2_$name.err').read()[-800:])
PY
}
EXTRA=""
run j18_base X=1
run j18_conv1 RDA_B200_LIB=$PWD/rda_planner_b200/librda_b200_conv1.so
EXTRA=""; run j18_nopolish RDA_B200_LIB=$PWD/rda_planner_b200/librda_b200_nopolish.so
EXTRA="--batch 1024"; run j18_b1024 X=1; run j18_b1024_conv1 RDA_B200_LIB=$PWD/rda_planner_b200/librda_b200_conv1.so; run j18_b1024_nopolish RDA_B200_LIB=$PWD/rda_planner_b200/librda_b200_nopolish.so
EXTRA="--batch 296"; run j18_b296 X=1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 48 --csv --log-file gpurun_out/s2_launches_18.csv env RDA_B200_SPLIT_MIN=1000000 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-probes > gpurun_out/s2_ncu_18.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/s2_launches_18.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    k = r[ki].split('(')[0][-40:]
    agg[k][0] += 1; agg[k][1] += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]): print(f'{k:42s} n={n:4d} total={t/1e3:9.1f} us  avg={t/n/1e3:8.1f} us')
PY
