#!/bin/bash
# one box, every BASELINE config back to back (the table of DESIGN 6), after the full GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02_pytest21.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest21.log; grep -n "^FAILED\|^E  " gpurun_out/r02_pytest21.log | head
B="python bench.py --no-e2e --no-cpu-baseline"
run() { name=$1; shift; $B "$@" > gpurun_out/r02_cfg_$name.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/r02_cfg_$name.json")); print("$name", "ms/step", round(d["ms_per_step"],3), "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "value %.3e"%d["value"], "recomputed", d["recomputed_rows_last_step"], "mhz", d["clocks"]["sm_mhz"], d["kernel"], "launches", d["gpu_launches"])
PY
}
run c2 --steps 20 --warmup 5
run c2_m01 --missing-rate 0.01 --steps 10 --warmup 3
run c3 --chained --missing-rate 0.25 --steps 5 --warmup 3
run c4 --phenotypes 128 --variants 200000 --steps 3 --warmup 2
run c5slice --samples 500000 --variants 800000 --steps 10 --warmup 3
run c1 --samples 1000 --variants 10000 --covariates-ignored 2>/dev/null || true
