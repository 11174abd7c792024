#!/bin/bash
cd "$(dirname "$0")/.."
export LRR_B200_LIB=$PWD/scratch/abl/tc4_abl.so
for cfg in "base:" "noMMA:LRR_ABL_BITS=2" "noSTTM:LRR_ABL_BITS=1" "noALU:LRR_ABL_BITS=8" "noPOPC:LRR_ABL_BITS=16" "noMMA_noSTTM:LRR_ABL_BITS=3" "noALU_noPOPC:LRR_ABL_BITS=24" "noB:LRR_ABL_BITS=4" "all_off:LRR_ABL_BITS=31"; do
  label=${cfg%%:*}; envs=${cfg#*:}
  out=$(env $envs python bench.py --kernel tc4 --missing-rate 0.01 --variants 250000 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$label: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("sweep_ms=%.3f"%d["roofline"]["kernel_ms"], "sm_mhz=%s"%d["clocks"]["sm_mhz"])' 2>&1 | tail -1)"
done
