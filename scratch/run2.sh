#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
tail -15 gpurun_out/r02_pytest2.log
python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_c2_b.json 2> gpurun_out/r02_bench_c2_b.err; tail -c 300 gpurun_out/r02_bench_c2_b.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_b.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_b.log 2>&1
grep -v "^==" gpurun_out/r02_launches_b.csv | awk -F'","' '{print $5, $(NF)}' | tail -30
python scratch/c1_e2e.py > gpurun_out/r02_c1_e2e_b.json 2> gpurun_out/r02_c1_e2e_b.err; cat gpurun_out/r02_c1_e2e_b.json; tail -c 300 gpurun_out/r02_c1_e2e_b.err
cut -c1-400 gpurun_out/r02_bench_c2_b.json
