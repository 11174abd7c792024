#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-e2e --no-cpu-baseline --phenotypes 128 --variants 200000 --steps 5 --warmup 3 > gpurun_out/r44_c4.json 2> gpurun_out/r44_c4.err
python bench.py --no-e2e --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r44_c2.json 2> gpurun_out/r44_c2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r44_launches_c4.csv python bench.py --no-e2e --no-cpu-baseline --phenotypes 128 --variants 200000 --steps 1 --warmup 1 > /dev/null 2>&1
python - <<'P'
import json,csv
for f in ('gpurun_out/r44_c4.json','gpurun_out/r44_c2.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, 'ms/step', round(d['ms_per_step'],2), 'sweep', r['kernel_ms'], 'non-sweep', round(d['ms_per_step']-r['kernel_ms'],2), d['clocks'].get('sm_mhz'))
rows=[r for r in csv.reader(open('gpurun_out/r44_launches_c4.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-12:]:
    if 'stats' in r[4]: print(r[4].split('(')[0][-40:], r[-1], r[-2])
P
