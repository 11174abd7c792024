#!/bin/bash
# sustained runs (about 2 s of sweeps each) with clock / power sampling: base vs ablations
cd "$(dirname "$0")/.."
for spec in "base:X=1" "noMMA:LRR_ABL_BITS=2" "noMMA_noB_noSTTM:LRR_ABL_BITS=7" "stream:LRR_ABL_STREAM=1"; do
  label=${spec%%:*}; envs=${spec#*:}
  out=$(env $envs timeout 300 python bench.py --kernel tc4 --variants 250000 --steps 400 --warmup 20 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$label: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("sweep_ms=%.3f"%d["roofline"]["kernel_ms"], "ms/step=%.3f"%d["ms_per_step"], d["clocks"])')"
done
