#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest12.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest12.log; grep -n "^FAILED\|^E  " gpurun_out/r02_pytest12.log | head -20
python bench_extra.py --which dense > gpurun_out/r02_bench_extra_dense.json 2> gpurun_out/r02_bench_extra_dense.err; python - <<'PY'
import json
for l in open("gpurun_out/r02_bench_extra_dense.json"):
    d=json.loads(l); print(d["config"]["workload"], d["config"]["missing_rate"], "ms", round(d["ms_per_step"],3), "entries/s %.3e"%d["value"], d["roofline"]["achieved"], d["roofline"]["frac"])
PY
tail -c 300 gpurun_out/r02_bench_extra_dense.err
python scratch/prof_e2e4.py > gpurun_out/r02_prof_e2e_b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_prof_e2e_b.json')); print([ (round(x['total_ms'],1), round(x['h2d_window_ms'],1)) for x in d])"
