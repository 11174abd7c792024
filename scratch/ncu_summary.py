"""Summarise an .ncu-rep (first kernel) into a text file of the metrics the design discussion uses."""
import csv
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct", "sm__inst_executed_pipe_tensor",
        "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread",
        "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct", "sm__issue_active.avg.pct", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__grid_size", "launch__cluster_size",
        "launch__block_size", "smsp__cycles_active.avg", "sm__inst_executed.sum", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "smsp__warp_issue_stalled", "launch__shared_mem_per_block_dynamic", "sm__pipe_shared_cycles_active.avg.pct"]
with open(out, "w") as f:
    f.write(title + "\n" + "=" * len(title) + "\n")
    for h, u, v in zip(hdr, units, vals):
        if any(w in h for w in want) and "realtime" not in h:
            f.write(f"{h:90s} {v} {u}\n")
print(open(out).read()[:300])
