"""world_size-2 gloo test (CPU) of the variant-sharding host logic: contiguous ranges, basis broadcast from rank 0,
row all-gather in rank order.  The per-shard compute is stood in by the numpy oracle (no GPU here); the result must
equal the single-process oracle run on the whole matrix."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hail_b200 import dist as hd
        from hail_b200.statgen import GroupBasis
        from oracle import linreg_oracle as O

        rng = np.random.default_rng(0)  # same data on every rank (as a shared file system would provide)
        N, M = 300, 101
        x = rng.integers(0, 3, size=(M, N)).astype(np.float64)
        x[rng.random(x.shape) < 0.1] = np.nan
        cov = np.column_stack([np.ones(N), rng.normal(size=(N, 3))])
        ys = rng.normal(size=(N, 2))
        ys[::17, 0] = np.nan

        bases = [GroupBasis(ys, cov, np.arange(N))] if rank == 0 else None
        bts = hd.broadcast_bases(bases, torch.device("cpu"))
        assert len(bts) == 1
        bt = bts[0]
        ref = GroupBasis(ys, cov, np.arange(N))
        assert (bt.n, bt.K, bt.P, bt.has_intercept) == (ref.n, ref.K, ref.P, int(ref.has_intercept))
        for t, f in zip(bt.tensors, hd.BASIS_FIELDS):
            assert np.array_equal(t.numpy(), getattr(ref, f)), f  # bit-identical after the broadcast

        lo, hi = hd.variant_range(rank, world, M)
        part = O.linreg_group(x[lo:hi], ys, cov)
        rows = torch.from_numpy(np.column_stack([part["sum_x"], part["beta"], part["p_value"]]))
        allrows = hd.gather_rows(rows).numpy()
        whole = O.linreg_group(x, ys, cov)
        want = np.column_stack([whole["sum_x"], whole["beta"], whole["p_value"]])
        assert allrows.shape == want.shape
        # (the stand-in oracle blocks rows by 16, so shard boundaries move its BLAS block shapes: last-bit differences)
        assert np.allclose(allrows, want, rtol=1e-10, atol=1e-13, equal_nan=True)
        assert np.array_equal(allrows[:, 0], want[:, 0], equal_nan=True)  # sum_x: exact, order preserved
        # Table.gather(): the sharded call's per-rank Tables concatenated in rank order (numeric fields as tensors,
        # object-valued row keys pickled)
        import hail_b200 as hb
        from collections import OrderedDict
        keys = np.array([("1", i + 1) for i in range(M)] + [None], dtype=object)[:-1]
        local = hb.Table(OrderedDict(locus=keys[lo:hi], n=np.full(hi - lo, 7, dtype=np.int32), sum_x=part["sum_x"],
                                     beta=part["beta"]), key=("locus",), n_rows=hi - lo)
        local.n_missing = np.arange(lo, hi, dtype=np.int32)
        full = local.gather()
        assert full.n_rows == M and list(full.locus) == list(keys) and np.array_equal(full.n_missing, np.arange(M))
        assert np.array_equal(full.sum_x, allrows[:, 0], equal_nan=True) and full.beta.shape == (M, 2)
        covered = [hd.variant_range(r, world, M) for r in range(world)]
        assert covered[0][0] == 0 and covered[-1][1] == M and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
        # the one path with a real exchange step: variant-sharded PCA all-reduces A' (A V) and the Gram matrices
        # (hail_b200/pca.py `_sharded`); the per-shard products are stood in by the numpy oracle
        from hail_b200 import pca as hpca
        from oracle import pca_oracle as PO
        a, keep = PO.hwe_normalize(x)                       # [m, N] with the GLOBAL variant count in the scaling
        kept_rows = np.nonzero(keep)[0]
        local = a[(kept_rows >= lo) & (kept_rows < hi)]
        V = np.random.default_rng(5).normal(size=(N, 4))
        T = local @ V
        W = hpca._allreduce(torch.from_numpy(local.T @ T), True).numpy()
        G = hpca._allreduce(torch.from_numpy(T.T @ T), True).numpy()
        assert np.allclose(W, a.T @ (a @ V), rtol=1e-11, atol=1e-12) and np.allclose(G, (a @ V).T @ (a @ V), rtol=1e-11)
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_variant_ranges_partition_everything():
    from hail_b200 import dist as hd

    for M in (0, 1, 7, 128, 1000003):
        for world in (1, 2, 3, 8):
            r = [hd.variant_range(k, world, M) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == M
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
