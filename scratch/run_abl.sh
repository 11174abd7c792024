#!/bin/bash
# run every ablation build on a 250k-variant slice; prints sweep kernel ms
cd "$(dirname "$0")/.."
run() { label=$1; shift; out=$(env "$@" python bench.py --variants 250000 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1); echo "$label: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["roofline"]["kernel_ms"], "ms", d["roofline"]["achieved"], "GB/s", d["clocks"]["sm_mhz"])' 2>&1 | tail -1)"; }
run base X=1
for v in no_bpanel no_geno no_mma no_sttm no_unpack mma_j2 no_bp_geno mma_n16; do run $v LRR_B200_LIB=$PWD/scratch/abl/$v.so; done
run cluster1 LRR_TC_CLUSTER=1
run cluster4 LRR_TC_CLUSTER=4
run gstages6 LRR_TC_GSTAGES=6
