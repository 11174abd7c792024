#!/bin/bash
# round 2, GPU run 1: the whole GPU suite, then the bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -40 gpurun_out/r02_pytest1.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_a.json 2> gpurun_out/r02_bench_c2_a.err; tail -c 600 gpurun_out/r02_bench_c2_a.err
python bench.py --chained --missing-rate 0.25 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_c3_a.json 2> gpurun_out/r02_bench_c3_a.err; tail -c 300 gpurun_out/r02_bench_c3_a.err
python bench.py --phenotypes 128 --variants 100000 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_c4_a.json 2> gpurun_out/r02_bench_c4_a.err; tail -c 300 gpurun_out/r02_bench_c4_a.err
python scratch/c1_e2e.py > gpurun_out/r02_c1_e2e_a.json 2> gpurun_out/r02_c1_e2e_a.err; tail -c 300 gpurun_out/r02_c1_e2e_a.err
for f in gpurun_out/r02_bench_c2_a.json gpurun_out/r02_bench_c3_a.json gpurun_out/r02_bench_c4_a.json gpurun_out/r02_c1_e2e_a.json; do echo "== $f"; cut -c1-1500 $f; done
