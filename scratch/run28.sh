#!/bin/bash
cd "$(dirname "$0")/.."
export LRR_B200_LIB=$PWD/scratch/abl/tc4_abl.so
for rep in 1 2; do
for cfg in "base:" "noB:LRR_ABL_BITS=4" "noB_noMMA:LRR_ABL_BITS=6"; do
  label=${cfg%%:*}; envs=${cfg#*:}
  out=$(env $envs timeout 200 python bench.py --kernel tc4 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$label: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); r=d["ranks"][0]; print("sweep_ms=%.3f"%d["roofline"]["kernel_ms"], "min/med/max", r["kernel_ms_min"], r["kernel_ms_median"], r["kernel_ms_max"], "sm_mhz=%s"%d["clocks"]["sm_mhz"])' 2>&1 | tail -1)"
done; done
