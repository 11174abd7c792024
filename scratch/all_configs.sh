#!/bin/bash
# one bench line per BASELINE.json config (C4 / C5 on slices that fit one GPU); prints a compact table
cd "$(dirname "$0")/.."
run() { label=$1; shift; out=$(timeout 600 python bench.py "$@" --no-cpu-baseline 2>/dev/null | tail -1); echo "$label: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); e=d.get("e2e",{}).get("value"); print(d["config"]["kernel"], "value=%.3e"%d["value"], "ms/step=%.3f"%d["ms_per_step"], "sweep_ms=%.3f"%d["roofline"]["kernel_ms"], "GB/s=%.0f"%d["roofline"]["achieved"], "frac=%.3f"%d["roofline"]["frac"], "e2e=%s"%("%.3e"%e if e else None))')"; }
run C1 --samples 1000 --variants 10000 --steps 20 --warmup 5 --e2e-variants 10000
run C2 --steps 5 --warmup 3 --no-e2e
run C3 --missing-rate 0.25 --chained --steps 3 --warmup 3 --no-e2e
run C4slice --phenotypes 128 --variants 100000 --steps 2 --warmup 3 --no-e2e
run C5slice --samples 500000 --variants 800000 --steps 3 --warmup 3 --no-e2e
