#!/bin/bash
export RDA_B200_NO_BUILD=1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest11.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest11.txt
tail -4 gpurun_out/r02_pytest11.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err; echo $name; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/r02_$name.json'))
    print(round(l['value']), l.get('roofline',{}).get('kernel_ms'), l.get('counters'), l['status_bits'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r02_$name.err').read()[-800:])
PY
}
EXTRA=""
run j11_base X=1
run j11_parts3 RDA_B200_SPLIT_PARTS=3
run j11_parts4 RDA_B200_SPLIT_PARTS=4
EXTRA="--batch 8192"; run j11_b8192 X=1
EXTRA="--batch 4096"; run j11_b4096 X=1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 40 --csv --log-file gpurun_out/r02_launches_11.csv env RDA_B200_SPLIT_MIN=1000000 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-probes > gpurun_out/r02_ncu_11.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02_launches_11.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    k = r[ki].split('(')[0][-40:]
    agg[k][0] += 1; agg[k][1] += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]): print(f'{k:42s} n={n:4d} total={t/1e3:9.1f} us  avg={t/n/1e3:8.1f} us')
PY
