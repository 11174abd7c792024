#!/bin/bash
# 8-GPU box: weak scaling of C2 (peer-copy vs NCCL gather), where the N=8 loss goes, strong scaling of C5, C4 on 8 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1
run() { # name nproc args...
  name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --gpus 1 --no-e2e --no-cpu-baseline "$@" > gpurun_out/r02_s_$name.json 2> gpurun_out/r02_s_$name.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --no-e2e --no-cpu-baseline "$@" > gpurun_out/r02_s_$name.json 2> gpurun_out/r02_s_$name.err; fi
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_s_$name.json"))
    ks=[r["kernel_ms_median"] for r in d["ranks"]]
    print("$name", "n", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "value %.4e"%d["value"], "kernel_ms min/max over ranks", min(ks), max(ks), "mhz", [r["sm_mhz"] for r in d["ranks"]], (d["multi_gpu"] or {}).get("gather"))
except Exception as e:
    print("$name FAILED", e); import subprocess; print(subprocess.run(["tail","-c","600","gpurun_out/r02_s_$name.err"],capture_output=True,text=True).stdout)
PY
}
run w1 1 --steps 20 --warmup 5
run w8_peer 8 --steps 20 --warmup 5 --gather peer
run w8_nccl 8 --steps 20 --warmup 5 --gather nccl
run w8_nogather 8 --steps 10 --warmup 3 --no-gather
run w8_nobcast 8 --steps 10 --warmup 3 --gather peer --no-broadcast
run w2_peer 2 --steps 20 --warmup 5 --gather peer
run w4_peer 4 --steps 20 --warmup 5 --gather peer
run c5_8 8 --strong --samples 500000 --variants 10000000 --steps 2 --warmup 1
run c5_4 4 --strong --samples 500000 --variants 10000000 --steps 2 --warmup 1
run c5_2 2 --strong --samples 500000 --variants 10000000 --steps 2 --warmup 1
run c4_8 8 --phenotypes 128 --variants 200000 --steps 3 --warmup 2
