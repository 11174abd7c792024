#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tolerance.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_pytest27.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest27.log; grep -n "^FAILED\|^E  " gpurun_out/r02_pytest27.log | head -20
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline"
run() { name=$1; shift; $B "$@" > gpurun_out/r02_cfg3_$name.json 2>/dev/null; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_cfg3_$name.json")); print("$name", "ms/step", round(d["ms_per_step"],3), "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "value %.3e"%d["value"], "recomputed", d["recomputed_rows_last_step"], "mhz", d["clocks"]["sm_mhz"], "launches", d["gpu_launches"])
except Exception as e: print("$name FAILED", e)
PY
}
run c2 --steps 10 --warmup 3
run c3 --chained --missing-rate 0.25 --steps 5 --warmup 3
run c3_1pct --chained --missing-rate 0.01 --steps 5 --warmup 3
run c4 --phenotypes 128 --variants 200000 --steps 3 --warmup 2
run c4_1pct --phenotypes 128 --variants 200000 --missing-rate 0.01 --steps 3 --warmup 2
