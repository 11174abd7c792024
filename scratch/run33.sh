#!/bin/bash
# centred dot products + slack: parity, then C2 / C4 / C3 / 1 % missing on the same box, old policy through the tuning build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r33_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r33_pytest.log
tail -3 gpurun_out/r33_pytest.log
C2="python bench.py --no-e2e --no-cpu-baseline --steps 20 --warmup 5"
$C2 > gpurun_out/r33_c2.json 2> gpurun_out/r33_c2.err
LRR_B200_LIB=$PWD/scratch/abl/tc4_abl.so LRR_TC4_SLACK=1 $C2 > gpurun_out/r33_c2_old96.json 2> gpurun_out/r33_c2_old96.err
$C2 > gpurun_out/r33_c2_b.json 2> gpurun_out/r33_c2_b.err
python bench.py --no-e2e --no-cpu-baseline --phenotypes 128 --variants 200000 --steps 5 --warmup 3 > gpurun_out/r33_c4.json 2> gpurun_out/r33_c4.err
python bench.py --no-e2e --no-cpu-baseline --chained --missing-rate 0.25 --steps 5 --warmup 3 > gpurun_out/r33_c3.json 2> gpurun_out/r33_c3.err
python bench.py --no-e2e --no-cpu-baseline --missing-rate 0.01 --steps 10 --warmup 3 > gpurun_out/r33_c2_1pct.json 2> gpurun_out/r33_c2_1pct.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r33_c*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']; t=r.get('tensor',{}); print(f, 'ms/step', round(d['ms_per_step'],2), 'sweep', r['kernel_ms'], 'hbm', r['frac'], 'tensor', t.get('frac'), t.get('sweep_launches'), t.get('mma_columns'), t.get('digit_columns_in_use'), 'recomputed', d.get('recomputed_rows_last_step'), d['clocks'].get('sm_mhz'))
P
