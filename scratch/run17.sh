#!/bin/bash
cd "$(dirname "$0")/.."
export LRR_B200_LIB=$PWD/scratch/abl/tc4_abl.so
for cfg in "3:224" "4:224" "5:224" "4:192" "5:192" "6:160" "5:176"; do
  b=${cfg%%:*}; w=${cfg##*:}
  LRR_TC4_BST=$b LRR_TC4_WIDE=$w python bench.py --no-e2e --no-cpu-baseline --phenotypes 128 --variants 200000 --steps 3 --warmup 2 > /tmp/o.json 2>/tmp/o.err
  python - <<PY
import json
try:
    d=json.load(open("/tmp/o.json")); print("bst=$b wide=$w", "sweep", d["roofline"]["kernel_ms"], "ms/step", round(d["ms_per_step"],2), d["gpu_launches"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("bst=$b wide=$w FAILED", e, open("/tmp/o.err").read()[-300:])
PY
done
