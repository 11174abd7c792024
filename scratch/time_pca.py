"""Wall time of hwe_normalized_pca on resident packed genotypes."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import hail_b200 as hb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mt = hb.balding_nichols_model(6, N, M, missing_rate=0.01, seed=5)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev, scores, _ = hb.hwe_normalized_pca(mt.GT, k=k)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"hwe_normalized_pca N={N} M={M} k={k}: {dt:.2f} s, {scores.n_iterations} iterations "
      f"({2 * scores.n_iterations} passes over the genotypes: {2 * scores.n_iterations * N * M / dt:.3e} genotypes/s), eigenvalues {np.round(ev, 3).tolist()}")
