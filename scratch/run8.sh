#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -q --timeout 900 -p no:cacheprovider -k "sharded or plugin or shapes or chained" > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest8.log
tail -8 gpurun_out/r02_pytest8.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-e2e --no-cpu-baseline --steps 20 --warmup 5"
$T --gather peer > gpurun_out/r02_bench_n2_peer2.json 2> gpurun_out/r02_bench_n2_peer2.err; tail -c 600 gpurun_out/r02_bench_n2_peer2.err
python bench.py --no-e2e --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_f.json 2> gpurun_out/r02_bench_n1_f.err
for f in n2_peer2 n1_f; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_${f}.json"))
    print("$f", "ms/step", round(d["ms_per_step"],3), "kernel_ms", d["roofline"]["kernel_ms"], "value %.3e"%d["value"], d["multi_gpu"], [ (r["kernel_ms_median"], r["region_ms_per_step"], r["sm_mhz"]) for r in d["ranks"]])
except Exception as e:
    print("$f", "FAILED", e)
PY
done
