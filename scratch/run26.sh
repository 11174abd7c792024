#!/bin/bash
cd "$(dirname "$0")/.."
for rep in 1 2 3; do
for v in 0 1; do
  out=$(LRR_B200_LIB=$PWD/scratch/abl/csa$v.so python bench.py --kernel tc4 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1)
  echo "csa=$v: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); r=d["ranks"][0]; print("sweep_ms=%.3f"%d["roofline"]["kernel_ms"], "min/med/max", r["kernel_ms_min"], r["kernel_ms_median"], r["kernel_ms_max"], "sm_mhz=%s"%d["clocks"]["sm_mhz"])' 2>&1 | tail -1)"
done; done
LRR_B200_LIB=$PWD/scratch/abl/csa1.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "full_sample or golden or fastlmm or mixed" 2>&1 | tail -2
