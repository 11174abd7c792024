#!/bin/bash
# wide-profile tail-digit slack (6 passes for C4?) + what C2 really runs (mma_columns)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tolerance.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r32_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r32_pytest.log
tail -3 gpurun_out/r32_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --phenotypes 128 --variants 200000 --steps 5 --warmup 3"
$B > gpurun_out/r32_c4_slack.json 2> gpurun_out/r32_c4_slack.err
LRR_B200_LIB=$PWD/scratch/abl/tc4_abl.so LRR_TC4_SLACK=1 LRR_TC4_WIDE_COVD=13 LRR_TC4_WIDE=224 $B > gpurun_out/r32_c4_old.json 2> gpurun_out/r32_c4_old.err
$B > gpurun_out/r32_c4_slack_b.json 2> gpurun_out/r32_c4_slack_b.err
C2="python bench.py --no-e2e --no-cpu-baseline --steps 20 --warmup 5"
$C2 > gpurun_out/r32_c2.json 2> gpurun_out/r32_c2.err
LRR_B200_LIB=$PWD/scratch/abl/tc4_abl.so LRR_TC4_SLACK=1.3 $C2 > gpurun_out/r32_c2_slack13.json 2> gpurun_out/r32_c2_slack13.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r32_c*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']; t=r.get('tensor',{}); print(f, 'ms/step', round(d['ms_per_step'],2), 'sweep', r['kernel_ms'], 'hbm', r['frac'], 'tensor', t.get('frac'), t.get('sweep_launches'), t.get('mma_columns'), t.get('digit_columns_in_use'), 'recomputed', d.get('recomputed_rows_last_step'), d['clocks'].get('sm_mhz'))
P
