import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hail_b200 as hb
from hail_b200 import _lib, bn
from hail_b200.statgen import GroupBasis, _push_groups, STAT_FIELDS

# ---- (1) compact vs dense: where do they differ?
rng = np.random.default_rng(41)
N, M, P = 1003, 150, 14
x = rng.uniform(0, 2, size=(M, N)); x[rng.random(x.shape) < 0.03] = np.nan; x[7] = np.nan
ys = rng.normal(size=(N, P)); ys[11, :] = np.nan
c = rng.normal(size=N)
for P_use in (1, 14):
    cd = hb.CompactDosage(x); deq = cd.to_dosage()
    cols = {**{f"y{i}": ys[:, i] for i in range(P_use)}, "c": c}
    cmt = hb.MatrixTable(cd, cols=cols); dmt = hb.MatrixTable(hb.DenseDosage(deq), cols=cols)
    hc = hb.linear_regression_rows(y=[cmt[f"y{i}"] for i in range(P_use)], x=cmt.x, covariates=[1.0, cmt.c])
    hd = hb.linear_regression_rows(y=[dmt[f"y{i}"] for i in range(P_use)], x=dmt.x, covariates=[1.0, dmt.c])
    for f in ("n", "sum_x", "y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
        a, b = np.asarray(hc[f], dtype=np.float64), np.asarray(hd[f], dtype=np.float64)
        bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
        print("P", P_use, f, int(bad.sum()), "maxrel", float(np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if bad.any() else 0.0, np.argwhere(bad)[:4].tolist())

# ---- (2) cost of one lrr_run on a 2560-variant block (the streaming loop's unit), host and device
N, Mb, K = 400_000, 2560, 10
dev = torch.device("cuda", 0)
pop, th, _ = bn.bn_parameters(3, N, Mb, seed=0)
gt = bn.bn_fill(hb.PackedGenotypes.empty(Mb, N, dev), pop, th, seed=0)
ctx = _lib.context(0)
rng = np.random.default_rng(1)
cov = np.column_stack([np.ones(N)] + [rng.standard_normal(N) for _ in range(K - 1)])
y = rng.standard_normal((N, 1))
_push_groups(ctx, N, [GroupBasis(y, cov, np.arange(N))])
o = {"n": torch.empty(Mb, dtype=torch.int32, device=dev), "n_missing": torch.empty(Mb, dtype=torch.int32, device=dev), "sum_x": torch.empty(Mb, dtype=torch.float64, device=dev)}
for f in STAT_FIELDS: o[f] = torch.empty((Mb, 1), dtype=torch.float64, device=dev)
arr = (_lib.GroupOut * 1)()
for k, v in o.items(): setattr(arr[0], k, v.data_ptr())
arr[0].log10_p = None
st = torch.cuda.Stream(dev)
def run(): ctx.check(ctx.lib.lrr_run(ctx.handle, gt.data.data_ptr(), gt.flags_ptr(), Mb, gt.stride, N, arr, 1, 0, st.cuda_stream))
for _ in range(3): run()
torch.cuda.synchronize()
for reps in (1, 20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(st)
    for _ in range(reps): run()
    e1.record(st); t_host = time.perf_counter() - t0
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    print("block run x%d: host issue %.3f ms per call, device %.3f ms per call, wall %.3f ms per call, kernel %s" % (reps, 1e3 * t_host / reps, e0.elapsed_time(e1) / reps, 1e3 * t_all / reps, ctx.last_kernel))
ctx.check(ctx.lib.lrr_set_timing(ctx.handle, 1)); run(); print("sweep ms", ctx.lib.lrr_last_sweep_ms(ctx.handle))
