"""Wall time of logistic_regression_rows (wald / lrt / firth) on resident packed genotypes: variants/s and sample-fits/s."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import hail_b200 as hb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mt = hb.balding_nichols_model(3, N, M, missing_rate=0.01, seed=5)
rng = np.random.default_rng(0)
cov = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(K - 1)])
y = (rng.random(N) < 1 / (1 + np.exp(-(0.3 * cov[:, 1] - 0.2)))).astype(np.float64)
mt = mt.annotate_cols(y=y, **{f"c{k}": cov[:, k] for k in range(K)})
for test in ("wald", "lrt", "firth", "score"):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ht = hb.logistic_regression_rows(test, mt.y, mt.GT.n_alt_alleles(), [mt[f"c{k}"] for k in range(K)])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    it = ht.fit["n_iterations"].mean() if test != "score" else 0
    print(f"logistic {test}: N={N} M={M} K={K}: {dt*1e3:.1f} ms (incl. host null fit), {M/dt:.0f} variants/s, "
          f"{N*M/dt:.3e} genotypes/s, mean iterations {it:.2f}")
