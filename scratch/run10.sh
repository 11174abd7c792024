#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scratch/prof_e2e4.py > gpurun_out/r02_prof_e2e.json 2> gpurun_out/r02_prof_e2e.err; cat gpurun_out/r02_prof_e2e.json; tail -c 400 gpurun_out/r02_prof_e2e.err
python bench.py --no-e2e --no-cpu-baseline --phenotypes 128 --variants 200000 --steps 3 --warmup 2 > gpurun_out/r02_bench_c4_g.json 2> gpurun_out/r02_bench_c4_g.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c4_g.json')); print('c4', d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'])"
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest10.log
