#!/bin/bash
# ncu evidence: launch list of the default bench, full captures of the C2 sweep, one wide C4 pass and one C3 pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline"
$B --steps 2 --warmup 1 > gpurun_out/r02_plain_c2.json 2> gpurun_out/r02_plain_c2.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_c2.csv $B --steps 2 --warmup 1 > gpurun_out/r02_ncu_l_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc4_sweep -s 1 -c 1 -o gpurun_out/r02_c2_full -f $B --steps 1 --warmup 1 > gpurun_out/r02_ncu_f_c2.log 2>&1
$B --phenotypes 128 --variants 200000 --steps 1 --warmup 1 > gpurun_out/r02_plain_c4.json 2> gpurun_out/r02_plain_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4.csv $B --phenotypes 128 --variants 200000 --steps 1 --warmup 1 > gpurun_out/r02_ncu_l_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc4_sweep -s 8 -c 1 -o gpurun_out/r02_c4_full -f $B --phenotypes 128 --variants 200000 --steps 1 --warmup 1 > gpurun_out/r02_ncu_f_c4.log 2>&1
$B --chained --missing-rate 0.25 --variants 250000 --steps 1 --warmup 1 > gpurun_out/r02_plain_c3.json 2> gpurun_out/r02_plain_c3.err
ncu --set full --clock-control none --import-source on -k regex:tc4_sweep -s 3 -c 1 -o gpurun_out/r02_c3_full -f $B --chained --missing-rate 0.25 --variants 250000 --steps 1 --warmup 1 > gpurun_out/r02_ncu_f_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
