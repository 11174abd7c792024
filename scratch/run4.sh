#!/bin/bash
# 2-GPU run: sharded product call (bit-for-bit vs one device), bench at N=2 with peer / nccl gathers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n2.txt 2>&1
python -m pytest tests/test_gpu_sharded.py tests/test_gpu_tolerance.py -m gpu -q --timeout 900 -p no:cacheprovider -k "sharded or refuses or c4" > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest4.log
tail -15 gpurun_out/r02_pytest4.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-e2e --no-cpu-baseline --steps 20 --warmup 5"
$T --gather peer > gpurun_out/r02_bench_n2_peer.json 2> gpurun_out/r02_bench_n2_peer.err; tail -c 1500 gpurun_out/r02_bench_n2_peer.err
$T --gather nccl > gpurun_out/r02_bench_n2_nccl.json 2> gpurun_out/r02_bench_n2_nccl.err; tail -c 300 gpurun_out/r02_bench_n2_nccl.err
python bench.py --no-e2e --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err
python bench.py --no-e2e --no-cpu-baseline --phenotypes 128 --variants 200000 --steps 3 --warmup 2 > gpurun_out/r02_bench_c4_d.json 2> gpurun_out/r02_bench_c4_d.err
for f in n2_peer n2_nccl n1_d c4_d; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_${f}.json"))
    print("$f", "ms/step", round(d["ms_per_step"],3), "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "value %.3e"%d["value"], d["multi_gpu"], [ (r["kernel_ms_median"], r["region_ms_per_step"], r["sm_mhz"]) for r in d["ranks"]])
except Exception as e:
    print("$f", "FAILED", e)
PY
done
