import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import hail_b200 as hb
N, M, k = 100000, 100000, 10
mt = hb.balding_nichols_model(6, N, M, missing_rate=0.01, seed=5)
ev, scores, _ = hb.hwe_normalized_pca(mt.GT, k=k, _max_iterations=3)
torch.cuda.synchronize(); t0 = time.perf_counter()
ev, scores, _ = hb.hwe_normalized_pca(mt.GT, k=k, _max_iterations=3)
torch.cuda.synchronize(); print("3 sweeps:", time.perf_counter() - t0)
