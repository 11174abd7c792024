#!/bin/bash
# sustained (100-step) sweep timing on a 250k-variant slice.  usage: sustain.sh label[:ENV=VAL,...] ...
cd "$(dirname "$0")/.."
for spec in "$@"; do
  label=${spec%%:*}; envs=""; [[ "$spec" == *:* ]] && envs=${spec#*:}
  out=$(env ${envs//,/ } timeout 120 python bench.py --kernel tc4 --variants 250000 --steps 100 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$label: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); c=d["clocks"]; print("sweep_ms=%.3f"%d["roofline"]["kernel_ms"], "GB/s=%.0f"%d["roofline"]["achieved"], "sm_mhz=%s"%c["sm_mhz"], "W=%s"%c.get("power_w"), "ncols_pad?")' 2>&1 | tail -1)"
done
