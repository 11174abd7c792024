import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import hail_b200 as hb
from hail_b200 import pca, statgen
N, M, k = 400000, 100000, 5
mt = hb.balding_nichols_model(6, N, M, missing_rate=0.01, seed=5)
hb.hwe_normalized_pca(mt.GT, k=k, _max_iterations=2)
# wall-time breakdown of one more call: wrap the two products
orig_run = pca._run_device
acc = {"sweep": 0.0, "n": 0}
def timed_run(*a, **kw):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig_run(*a, **kw)
    torch.cuda.synchronize(); acc["sweep"] += time.perf_counter() - t0; acc["n"] += 1
    return r
pca._run_device = timed_run
torch.cuda.synchronize(); t0 = time.perf_counter()
ev, scores, _ = hb.hwe_normalized_pca(mt.GT, k=k, _max_iterations=3)
torch.cuda.synchronize(); tot = time.perf_counter() - t0
print(f"total {tot:.3f} s, {acc['n']} sweeps (lrr_add_group + lrr_run + epilogue) {acc['sweep']:.3f} s")
