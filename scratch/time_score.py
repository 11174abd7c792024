import time, numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hail_b200 as hb
from hail_b200 import bn
N, M, K = 400000, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 10
pop, th, _ = bn.bn_parameters(3, N, M, missing_rate=0.0, seed=0)
gt = bn.bn_fill(hb.PackedGenotypes.empty(M, N, 0), pop, th, seed=0)
rng = np.random.default_rng(1)
cov = np.column_stack([np.ones(N)] + [rng.standard_normal(N) for _ in range(K - 1)])
y = (rng.random(N) < 0.3).astype(np.float64)
mt = hb.MatrixTable(gt, cols={"y": y, **{f"c{i}": cov[:, i] for i in range(1, K)}})
covs = [1.0] + [mt[f"c{i}"] for i in range(1, K)]
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    ht = hb.logistic_regression_rows("score", mt.y, mt.GT.n_alt_alleles(), covs)
    torch.cuda.synchronize(); dt = time.time() - t
    print("logistic score: %d variants x %d samples in %.3f s -> %.3e genotypes/s (incl. host null fit)" % (M, N, dt, M * N / dt))
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    ht = hb.linear_regression_rows(mt.y, mt.GT.n_alt_alleles(), covs, weights=mt.c1 * 0 + 1.5 if False else None, _kernel="fp64")
    torch.cuda.synchronize(); dt = time.time() - t
    print("linear fp64 kernel: %.3f s -> %.3e genotypes/s" % (dt, M * N / dt))
