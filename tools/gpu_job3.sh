#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest3.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest3.txt
tail -3 gpurun_out/r02_pytest3.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-probes"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err; echo $name; python - <<PY
import json
try:
    l=json.load(open('gpurun_out/r02_$name.json'))
    print(round(l['value']), l['roofline']['kernel_ms'], l['counters'], l['status_bits'], l['gpu_launches'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r02_$name.err').read()[-600:])
PY
}
EXTRA=""
run sb_auto X=1
run sb_off RDA_B200_SU_BATCHED=0
run sb_nosplit RDA_B200_SPLIT_MIN=100000
EXTRA="--batch 4096"
run sb_b4096_on RDA_B200_SU_BATCHED=1
run sb_b4096_off RDA_B200_SU_BATCHED=0
EXTRA="--batch 1024"
run sb_b1024_on RDA_B200_SU_BATCHED=1
# launch list of the batched path (first 400 launches after warm-up are enough to see the kernel mix)
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 -c 600 --csv --log-file gpurun_out/r02_launches_sb.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-probes > gpurun_out/r02_ncu_sb.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02_launches_sb.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    k = r[ki].split('(')[0][-40:]
    agg[k][0] += 1; agg[k][1] += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]): print(f'{k:42s} n={n:4d} total={t/1e3:9.1f} us  avg={t/n/1e3:8.1f} us')
PY
