#!/bin/bash
# quick perf probe: sweep kernel ms on a 250k-variant slice.  usage: quick.sh label[:ENV=VAL[,ENV=VAL]][@kernel] ...
cd "$(dirname "$0")/.."
for spec in "$@"; do
  kern=auto; [[ "$spec" == *@* ]] && { kern=${spec##*@}; spec=${spec%@*}; }
  label=${spec%%:*}; envs=""; [[ "$spec" == *:* ]] && envs=${spec#*:}
  out=$(env ${envs//,/ } timeout 200 python bench.py --kernel $kern --variants 250000 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$label ($kern): $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["config"]["kernel"], d["roofline"]["kernel_ms"], "ms", d["roofline"]["achieved"], "GB/s", d["clocks"]["sm_mhz"])' 2>&1 | tail -1)"
done
