#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest6.log
tail -8 gpurun_out/r02_pytest6.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke6.log 2>&1; echo "smoke rc=$?"; tail -12 gpurun_out/r02_smoke6.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_e.json 2> gpurun_out/r02_bench_c2_e.err; cut -c1-2500 gpurun_out/r02_bench_c2_e.json
