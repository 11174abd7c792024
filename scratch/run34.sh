#!/bin/bash
# after the centred bound: full GPU suite, then ncu evidence of the default bench (launch list + full C2 sweep), one 240-column C4 pass, one C3 plane sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r34_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r34_pytest.log
tail -3 gpurun_out/r34_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline"
$B --steps 2 --warmup 1 > gpurun_out/r34_plain_c2.json 2> gpurun_out/r34_plain_c2.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r34_launches_c2.csv $B --steps 2 --warmup 1 > gpurun_out/r34_ncu_l_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc4_sweep -s 1 -c 1 -o gpurun_out/r34_c2_full -f $B --steps 1 --warmup 1 > gpurun_out/r34_ncu_f_c2.log 2>&1
$B --phenotypes 128 --variants 200000 --steps 1 --warmup 1 > gpurun_out/r34_plain_c4.json 2> gpurun_out/r34_plain_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r34_launches_c4.csv $B --phenotypes 128 --variants 200000 --steps 1 --warmup 1 > gpurun_out/r34_ncu_l_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc4_sweep -s 7 -c 1 -o gpurun_out/r34_c4_full -f $B --phenotypes 128 --variants 200000 --steps 1 --warmup 1 > gpurun_out/r34_ncu_f_c4.log 2>&1
$B --chained --missing-rate 0.25 --variants 250000 --steps 1 --warmup 1 > gpurun_out/r34_plain_c3.json 2> gpurun_out/r34_plain_c3.err
ncu --set full --clock-control none --import-source on -k regex:tc4_sweep -s 2 -c 1 -o gpurun_out/r34_c3_full -f $B --chained --missing-rate 0.25 --variants 250000 --steps 1 --warmup 1 > gpurun_out/r34_ncu_f_c3.log 2>&1
ls -la gpurun_out/r34*.ncu-rep
