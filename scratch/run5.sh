#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LRR_B200_LIB=$PWD/scratch/abl/tc4_abl.so
for w in 112 128 160 192 224; do
  LRR_TC4_WIDE=$w python bench.py --no-e2e --no-cpu-baseline --phenotypes 128 --variants 200000 --steps 3 --warmup 2 > gpurun_out/r02_c4_wide_$w.json 2> gpurun_out/r02_c4_wide_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_c4_wide_$w.json"))
    print("wide=$w", "ms/step", round(d["ms_per_step"],3), "kernel_ms", d["roofline"]["kernel_ms"], "launches", d["gpu_launches"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("wide=$w FAILED", e)
PY
done
