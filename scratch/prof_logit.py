import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import hail_b200 as hb
N, M, K = 400000, 1184, 10
mt = hb.balding_nichols_model(3, N, M, missing_rate=0.01, seed=5)
rng = np.random.default_rng(0)
cov = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(K - 1)])
y = (rng.random(N) < 1 / (1 + np.exp(-(0.3 * cov[:, 1] - 0.2)))).astype(np.float64)
mt = mt.annotate_cols(y=y, **{f"c{k}": cov[:, k] for k in range(K)})
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ht = hb.logistic_regression_rows("wald", mt.y, mt.GT.n_alt_alleles(), [mt[f"c{k}"] for k in range(K)])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"wald {M} variants x {N}: {dt*1e3:.1f} ms, mean iterations {ht.fit['n_iterations'].mean():.2f}")
