import time, numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hail_b200 as hb
from hail_b200 import _lib
from hail_b200.statgen import GroupBasis
N = 400000; Me = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
print("cpus", os.cpu_count())
rng = np.random.default_rng(0)
cov = np.column_stack([np.ones(N)] + [rng.standard_normal(N) for _ in range(9)]); y = rng.standard_normal((N, 1))
for r in range(3):
    t = time.time(); b = GroupBasis(y, cov, np.arange(N)); print("GroupBasis %.3f s" % (time.time() - t))
t = time.time(); h = torch.empty((Me, (N + 3) // 4), dtype=torch.uint8, pin_memory=True); print("pin alloc %.2f s" % (time.time() - t))
h.random_(0, 255)
d = torch.empty_like(h, device="cuda")
for r in range(3):
    torch.cuda.synchronize(); t = time.time(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.time() - t
    print("H2D %.1f GB in %.3f s = %.1f GB/s" % (h.numel() / 1e9, dt, h.numel() / 1e9 / dt))
del d
col = {"y": y[:, 0], **{f"c{i}": cov[:, i] for i in range(1, 10)}}
def step():
    g = hb.HostBedGenotypes(h, N, 0)
    mt = hb.MatrixTable(g, cols=col)
    return hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0] + [mt[f"c{i}"] for i in range(1, 10)])
for r in range(3):
    t = time.time(); ht = step(); print("e2e step %.3f s -> %.3e genotypes/s" % (time.time() - t, Me * N / (time.time() - t)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); step(); pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
