#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest3.log
tail -12 gpurun_out/r02_pytest3.log
B="python bench.py --no-e2e --no-cpu-baseline"
$B --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_c.json 2> gpurun_out/r02_bench_c2_c.err; tail -c 300 gpurun_out/r02_bench_c2_c.err
$B --chained --missing-rate 0.25 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_c.json 2> gpurun_out/r02_bench_c3_c.err; tail -c 300 gpurun_out/r02_bench_c3_c.err
$B --missing-rate 0.01 --steps 5 --warmup 3 > gpurun_out/r02_bench_m01_c.json 2> gpurun_out/r02_bench_m01_c.err; tail -c 300 gpurun_out/r02_bench_m01_c.err
$B --phenotypes 128 --variants 100000 --steps 3 --warmup 2 > gpurun_out/r02_bench_c4_c.json 2> gpurun_out/r02_bench_c4_c.err; tail -c 300 gpurun_out/r02_bench_c4_c.err
for f in c2 c3 m01 c4; do python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_${f}_c.json"))
print("$f", "ms/step", round(d["ms_per_step"],3), "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "value %.3e"%d["value"], "recomputed", d["recomputed_rows_last_step"], d["clocks"]["sm_mhz"], d["kernel"], d["gpu_launches"])
PY
done
