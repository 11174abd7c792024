"""cProfile of one end-to-end public call on host-resident .bed rows: where the ~25 ms outside the streamed phases go."""
import cProfile, io, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hail_b200 as hb
from hail_b200 import _lib, bn, statgen

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
y = rng.standard_normal(N)
col = {"y": y, **{f"c{i}": cov[:, i] for i in range(1, K)}}

def step():
    g = hb.HostBedGenotypes(h_bed, N, dev)
    mt = hb.MatrixTable(g, cols=col)
    return hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0] + [mt[f"c{i}"] for i in range(1, K)])

for _ in range(3):
    step()
ts = []
for _ in range(8):
    t0 = time.perf_counter(); ht = step(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print("rep %.1f ms" % (1e3 * ts[-1]), {k: round(v, 1) for k, v in statgen.LAST_STREAM_PHASES.items()})
pr = cProfile.Profile(); pr.enable(); ht = step(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
