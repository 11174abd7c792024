#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_tolerance.py tests/test_gpu_impex.py -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02_pytest20.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest20.log; grep -n "^FAILED\|^E  " gpurun_out/r02_pytest20.log | head
B="python bench.py --no-e2e --no-cpu-baseline"
for cfg in "m01:--missing-rate 0.01 --steps 10 --warmup 3" "c3:--chained --missing-rate 0.25 --steps 5 --warmup 3" "c2:--steps 20 --warmup 5"; do
  name=${cfg%%:*}; args=${cfg#*:}
  $B $args > gpurun_out/r02_bench_${name}_h.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_${name}_h.json")); print("$name", "ms/step", round(d["ms_per_step"],3), "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], d["clocks"]["sm_mhz"])
PY
done
