#!/bin/bash
export RDA_B200_NO_BUILD=1
python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest15.txt 2>&1; tail -4 gpurun_out/r02_pytest15.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run() { name=$1; shift; env "$@" $B > gpurun_out/r02_j15_$name.json 2>gpurun_out/r02_j15_$name.err; python -c "
import json; l=json.load(open('gpurun_out/r02_j15_$name.json')); print('$name', round(l['value']), l['ms_per_step'], l['roofline']['kernel_ms'], l.get('counters',{}).get('cells_slow'))" || tail -3 gpurun_out/r02_j15_$name.err; }
run base X=1
run cpw32 RDA_B200_SLOW_CPW=32 RDA_B200_SLOW_ADAPT=0
run ctas8 RDA_B200_SLOW_CTAS=8
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_j15_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-probes > gpurun_out/r02_j15_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_j15_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki].split('(')[0][:40]].append(float(r[vi].replace(',','')))
    except Exception: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print(k, len(v), 'avg us', round(sum(v)/len(v)/1e3,1), 'max', round(max(v)/1e3,1))
PY
