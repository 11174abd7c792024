import time, numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
import hail_b200 as hb
from hail_b200 import _lib, bn
from hail_b200.statgen import GroupBasis
N=400000; Me=32768
rng=np.random.default_rng(0)
cov=np.column_stack([np.ones(N)]+[rng.standard_normal(N) for _ in range(9)]); y=rng.standard_normal((N,1))
t=time.time(); b=GroupBasis(y,cov,np.arange(N)); print("GroupBasis", time.time()-t)
t=time.time(); b=GroupBasis(y,cov,np.arange(N)); print("GroupBasis again", time.time()-t)
h=torch.randint(0,255,(Me,(N+3)//4),dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
for rep in range(3):
    t=time.time(); g=hb.PackedGenotypes.from_bed_rows(h,N,0,chunk_variants=4096); torch.cuda.synchronize(); print("from_bed_rows", time.time()-t)
mt=hb.MatrixTable(g, cols={"y":y[:,0], **{f"c{i}":cov[:,i] for i in range(1,10)}})
for rep in range(3):
    t=time.time(); ht=hb.linear_regression_rows(y=mt.y,x=mt.GT.n_alt_alleles(),covariates=[1.0]+[mt[f"c{i}"] for i in range(1,10)]); print("linreg call", time.time()-t)
import cProfile, pstats
pr=cProfile.Profile(); pr.enable()
ht=hb.linear_regression_rows(y=mt.y,x=mt.GT.n_alt_alleles(),covariates=[1.0]+[mt[f"c{i}"] for i in range(1,10)])
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
