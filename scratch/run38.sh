#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r38_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r38_pytest.log
tail -3 gpurun_out/r38_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r38_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r38_ref.json 2> gpurun_out/r38_ref.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r38_bench.json 2> gpurun_out/r38_bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r38_ref.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/r38_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['tensor']['mma_columns'], d['e2e']['value'], d['e2e']['rep_seconds'], d['cpu_baseline']['value'], d['clocks'])
P
