#!/bin/bash
# 8-GPU box: weak scaling with the peer-copy broadcast + gather (default transports), same-box N=1, default bench line at N=8 incl. e2e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@" > gpurun_out/r02_t_$name.json 2> gpurun_out/r02_t_$name.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n "$@" > gpurun_out/r02_t_$name.json 2> gpurun_out/r02_t_$name.err; fi
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_t_$name.json"))
    ks=[r["kernel_ms_median"] for r in d["ranks"]]
    e=d.get("e2e") or {}
    print("$name", "n", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "value %.4e"%d["value"], "kernel_ms", min(ks), max(ks), "e2e %.3e"%e.get("value",0), e.get("h2d_gbps_in_call"), e.get("h2d_gbps_link_alone"), d.get("multi_gpu"))
except Exception as ex:
    print("$name FAILED", ex); import subprocess; print(subprocess.run(["tail","-c","800","gpurun_out/r02_t_$name.err"],capture_output=True,text=True).stdout)
PY
}
run n1 1 --steps 20 --warmup 5 --no-cpu-baseline
run n8 8 --steps 20 --warmup 5
run n4 4 --steps 20 --warmup 5 --no-e2e
run n2 2 --steps 20 --warmup 5 --no-e2e
