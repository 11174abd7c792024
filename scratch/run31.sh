#!/bin/bash
# 240-column wide passes + 12-digit wide-profile covariate columns: parity, then C4 A/B on the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r31_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r31_pytest.log
tail -3 gpurun_out/r31_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --phenotypes 128 --variants 200000 --steps 5 --warmup 3"
$B > gpurun_out/r31_c4_240.json 2> gpurun_out/r31_c4_240.err
LRR_B200_LIB=$PWD/scratch/abl/tc4_abl.so LRR_TC4_WIDE=224 LRR_TC4_WIDE_COVD=13 $B > gpurun_out/r31_c4_224_tuning.json 2> gpurun_out/r31_c4_224_tuning.err
LRR_B200_LIB=$PWD/scratch/abl/tc4_abl.so $B > gpurun_out/r31_c4_240_tuning.json 2> gpurun_out/r31_c4_240_tuning.err
$B > gpurun_out/r31_c4_240_b.json 2> gpurun_out/r31_c4_240_b.err
python bench.py --no-e2e --no-cpu-baseline --chained --missing-rate 0.25 --steps 5 --warmup 3 > gpurun_out/r31_c3.json 2> gpurun_out/r31_c3.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r31_c*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']; print(f, 'ms/step', round(d['ms_per_step'],2), 'sweep', r['kernel_ms'], 'hbm', r['frac'], 'tensor', r.get('tensor',{}).get('frac'), r.get('tensor',{}).get('sweep_launches'), r.get('tensor',{}).get('mma_columns'), 'recomputed', d.get('recomputed_rows_last_step'), d['clocks'].get('sm_mhz'))
P
