#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02_pytest23.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest23.log; grep -n "^FAILED\|^E  " gpurun_out/r02_pytest23.log | head
B="python bench.py --no-e2e --no-cpu-baseline"
run() { name=$1; shift; $B "$@" > gpurun_out/r02_cfg2_$name.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/r02_cfg2_$name.json")); print("$name", "ms/step", round(d["ms_per_step"],3), "kernel_ms", d["roofline"]["kernel_ms"], "non-sweep", round(d["ms_per_step"]-d["roofline"]["kernel_ms"],3), "frac", d["roofline"]["frac"], "mhz", d["clocks"]["sm_mhz"], "launches", d["gpu_launches"])
PY
}
run c2 --steps 20 --warmup 5
run c4 --phenotypes 128 --variants 200000 --steps 3 --warmup 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_c4_b.csv $B --phenotypes 128 --variants 200000 --steps 1 --warmup 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_launches_c4_b.csv | python -c "
import csv,sys,collections
rows=list(csv.reader(sys.stdin))[1:]
agg=collections.OrderedDict()
for r in rows[-24:]:
    k=r[4][:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[-1])/1e6
for k,(n,t) in agg.items(): print(f'{k:62s} x{n:4d} {t:9.3f} ms')
"
