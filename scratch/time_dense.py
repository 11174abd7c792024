"""Device-event timing of the dense float64 sweep (lrr_run_dense) at one size: GB/s of the 8-byte entries."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import hail_b200 as hb
from hail_b200 import statgen, _lib
from hail_b200.statgen import GroupBasis

N = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
miss = float(sys.argv[4]) if len(sys.argv) > 4 else 0.01
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
x = torch.rand((M, N), device=dev, dtype=torch.float64, generator=g) * 2
x[torch.rand((M, N), device=dev, generator=g) < miss] = float("nan")
rng = np.random.default_rng(0)
cov = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(K - 1)])
y = rng.normal(size=(N, 1))
dd = hb.DenseDosage(x, 0)
bases = [GroupBasis(y, cov, np.arange(N), None)]
ctx = _lib.context(0)
statgen._push_groups(ctx, N, bases)
o = {"n": torch.empty(M, dtype=torch.int32, device=dev), "n_missing": torch.empty(M, dtype=torch.int32, device=dev),
     "sum_x": torch.empty(M, dtype=torch.float64, device=dev)}
for f in statgen.STAT_FIELDS:
    o[f] = torch.empty((M, 1), dtype=torch.float64, device=dev)
arr = (_lib.GroupOut * 1)()
for k, v in o.items():
    setattr(arr[0], k, v.data_ptr())
arr[0].log10_p = None
stream = torch.cuda.current_stream(dev).cuda_stream
def run():
    ctx.check(ctx.lib.lrr_run_dense(ctx.handle, dd.data.data_ptr(), M, N, N, arr, 1, stream))
for it in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for it in range(reps):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"dense N={N} M={M} K={K} miss={miss}: {ms:.3f} ms/run (sweep + imputation + statistics), {M*N*8/ms/1e6:.0f} GB/s of entries, {M*N/ms*1e3:.3e} entries/s")
