"""torchrun --nproc-per-node 2: variant-sharded hwe_normalized_pca (NCCL all-reduce of A'T and T'T) against the
single-GPU run on the whole matrix."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
import hail_b200 as hb
from hail_b200 import dist as hd

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.cuda.current_device()
dist.init_process_group("nccl")
N, M, k = 3000, 20000, 4
full = hb.balding_nichols_model(5, N, M, missing_rate=0.01, seed=9, device=dev)
lo, hi = hd.variant_range(rank, world, M)
shard = hb.MatrixTable(full.genotypes.rows(lo, hi), rows={"v": np.arange(lo, hi)}, cols={"s": np.arange(N)}, row_key=("v",), col_key=("s",))
ev_s, sc_s, ld_s = hb.hwe_normalized_pca(shard.GT, k=k, compute_loadings=True, _sharded=True)
whole = hb.MatrixTable(full.genotypes, rows={"v": np.arange(M)}, cols={"s": np.arange(N)}, row_key=("v",), col_key=("s",))
ev, sc, ld = hb.hwe_normalized_pca(whole.GT, k=k, compute_loadings=True)
ok_ev = np.allclose(ev_s, ev, rtol=1e-7)
sgn = np.sign(np.sum(sc_s.scores * sc.scores, axis=0))
ok_sc = np.allclose(sc_s.scores * sgn, sc.scores, rtol=1e-5, atol=1e-7)
mine = (ld.v >= lo) & (ld.v < hi)
ok_ld = ld_s.count() == int(mine.sum()) and np.allclose(ld_s.loadings * sgn, ld.loadings[mine], rtol=1e-5, atol=1e-8)
print(f"rank {rank}: sharded eigenvalues {np.round(ev_s, 4).tolist()} iterations {sc_s.n_iterations}; equal to the single-GPU run: "
      f"eigenvalues {ok_ev}, scores {ok_sc}, loadings {ok_ld}")
assert ok_ev and ok_sc and ok_ld
dist.barrier(); dist.destroy_process_group()
