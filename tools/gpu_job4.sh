#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest4.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest4.txt
grep -E "^it |passed|failed|rc " gpurun_out/r02_pytest4.txt | tail -14
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err
python - <<'PY'
import json
try:
    l=json.load(open('gpurun_out/r02_bench4.json'))
    print(round(l['value']), round(l['e2e']['value']), l['roofline']['kernel_ms'], l['status_bits'], l['cpu_baseline'], l.get('single_instance'), l.get('early_stop'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r02_bench4.err').read()[-800:])
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ref4.json 2> gpurun_out/r02_ref4.err; python -c "
import json; l=json.load(open('gpurun_out/r02_ref4.json')); print(round(l['value']), l['cpu_baseline'])"
