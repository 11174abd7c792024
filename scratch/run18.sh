#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest18.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest18.log; grep -n "^FAILED\|^E  " gpurun_out/r02_pytest18.log | head -20
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke18.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default_g.json 2> gpurun_out/r02_bench_default_g.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_default_g.json"))
e=d["e2e"]
print("C2 default: ms/step", round(d["ms_per_step"],3), "kernel", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "value %.3e"%d["value"], "mhz", d["clocks"]["sm_mhz"],
      "| e2e %.3e"%e["value"], e["rep_seconds"], "in-call", e["h2d_gbps_in_call"], "alone", e["h2d_gbps_link_alone"], "| cpu %.3e"%d["cpu_baseline"]["value"])
PY
python bench_extra.py --which logistic > gpurun_out/r02_bench_extra_logistic.json 2> gpurun_out/r02_bench_extra_logistic.err; python - <<'PY'
import json
for l in open("gpurun_out/r02_bench_extra_logistic.json"):
    d=json.loads(l); print(d["config"]["workload"], "%.3e genotypes/s"%d["value"], round(d["variants_per_s"]), "variants/s | cpu %.3e"%d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"])
PY
tail -c 300 gpurun_out/r02_bench_extra_logistic.err
