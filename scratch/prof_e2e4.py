"""Where the end-to-end call spends its time: stream_begin | host prologue | add_group + stream_run | table."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hail_b200 as hb
from hail_b200 import _lib, bn
from hail_b200.statgen import GroupBasis, _HostStream

N, Me, K = 400_000, 131072, 10
dev = torch.device("cuda", 0)
pop, th, _ = bn.bn_parameters(3, N, Me, seed=0)
gt = bn.bn_fill(hb.PackedGenotypes.empty(Me, N, dev), pop, th, seed=0)
ctx = _lib.context(0)
bed_stride = (N + 3) // 4
d_bed = torch.empty((Me, bed_stride), dtype=torch.uint8, device=dev)
ctx.check(ctx.lib.lrr_unpack_bed(ctx.handle, gt.data.data_ptr(), gt.stride, Me, N, d_bed.data_ptr(), bed_stride, None))
h_bed = torch.empty((Me, bed_stride), dtype=torch.uint8, pin_memory=True)
h_bed.copy_(d_bed); del d_bed, gt
torch.cuda.synchronize()
rng = np.random.default_rng(1)
cov = np.column_stack([np.ones(N)] + [rng.standard_normal(N) for _ in range(K - 1)])
y = rng.standard_normal((N, 1))
g = hb.HostBedGenotypes(h_bed, N, dev)
out = []
blocks = [int(v) for v in os.environ.get("BLOCKS", "0").split(",")]
ctx.check(ctx.lib.lrr_set_timing(ctx.handle, 1))
import ctypes
for rep in range(3 * len(blocks) + 2):
    blk = blocks[max(rep - 2, 0) // 3]
    t0 = time.perf_counter()
    st = _HostStream(g, block_variants=blk)
    t1 = time.perf_counter()
    bases = [GroupBasis(y, cov, np.arange(N))]
    t2 = time.perf_counter()
    host = st.run(bases)
    t3 = time.perf_counter()
    st.close()
    t4 = time.perf_counter()
    buf = (ctypes.c_float * 12288)()
    nt = ctx.lib.lrr_last_stream_timeline(ctx.handle, buf, 12288)
    tl = np.array(buf[:min(nt, 12288)]).reshape(-1, 3)
    if rep == 3:
        print("timeline (copy done, sweep started, statistics done) every 5th block:\n", np.round(tl[::5], 1))
    out.append({"block": blk, "last_block_done_ms": float(tl[-1, 2]) if len(tl) else -1.0, "first_sweep_start_ms": float(tl[0, 1]) if len(tl) else -1.0, "stream_begin_ms": 1e3 * (t1 - t0), "prologue_ms": 1e3 * (t2 - t1), "run_ms": 1e3 * (t3 - t2), "close_ms": 1e3 * (t4 - t3),
                "total_ms": 1e3 * (t4 - t0), "h2d_window_ms": float(ctx.lib.lrr_last_stream_h2d_ms(ctx.handle))})
print("\n".join(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in o.items()}) for o in out))
