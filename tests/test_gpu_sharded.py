"""Sharding INSIDE the drop-in call (LinearRegression.scala:95, :274 one task per partition; :74-78, :257 one broadcast
per call): two processes, one GPU each, every rank passes its contiguous range of the rows; the gathered Table must
equal the single-device run of the whole dataset BIT FOR BIT.  Needs two GPUs (skipped on a one-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dataset(hb, rank=None, world=None):
    """The same seeded dataset everywhere; `rank` selects that rank's contiguous row range."""
    from hail_b200 import bn
    from hail_b200 import dist as hd
    N, M, K = 60_000, 12_000, 6   # every shard is still large enough for AUTO to take the tensor-core sweep
    lo, hi = (0, M) if rank is None else hd.variant_range(rank, world, M)
    pop, th, _ = bn.bn_parameters(3, N, M, missing_rate=0.02, seed=9)
    gt = bn.bn_fill(hb.PackedGenotypes.empty(hi - lo, N, torch.device("cuda", torch.cuda.current_device())), pop,
                    th[lo:hi], seed=9, first_variant=lo)
    rng = np.random.default_rng(3)
    cov = rng.standard_normal((N, K - 1))
    y1 = rng.standard_normal(N)
    y2 = rng.standard_normal(N)
    y1[rng.random(N) < 0.05] = np.nan
    y2[rng.random(N) < 0.1] = np.nan
    rows = {"locus": np.array([("1", i + 1) for i in range(lo, hi)] + [None], dtype=object)[:-1],
            "rsid": np.array([f"rs{i}" for i in range(lo, hi)], dtype=object)}
    mt = hb.MatrixTable(gt, rows=rows, cols={"y1": y1, "y2": y2, **{f"c{i}": cov[:, i] for i in range(K - 1)}}, row_key=("locus",))
    return mt, K


def _call(hb, mt, K, **kw):
    return hb.linear_regression_rows(y=[[mt.y1], [mt.y2]], x=mt.GT.n_alt_alleles(),
                                     covariates=[1.0] + [mt[f"c{i}"] for i in range(K - 1)], pass_through=["rsid"], **kw)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import hail_b200 as hb
        mt, K = _dataset(hb, rank, world)
        for kernel in ("auto", "tc4", "fp64"):
            # the basis message through NCCL, and (tc4) pulled through symmetric memory by the copy engines
            local = _call(hb, mt, K, _kernel=kernel, _sharded="peer" if kernel == "tc4" else True)
            assert local.n_rows == mt.count_rows()
            full = local.gather()
            if rank == 0:
                np.savez(os.path.join(tmp, f"sharded_{kernel}.npz"), n=full.n, sum_x=full.sum_x, n_missing0=full.n_missing[0],
                         rsid=np.asarray(full.rsid, dtype=str), **{f"{f}{g}": full[f][g] for g in range(2)
                                                                  for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value")})
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_call_equals_single_device_bit_for_bit(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
    import hail_b200 as hb
    torch.cuda.set_device(0)
    mt, K = _dataset(hb)
    for kernel in ("auto", "tc4", "fp64"):
        one = _call(hb, mt, K, _kernel=kernel)
        z = np.load(tmp_path / f"sharded_{kernel}.npz")
        assert np.array_equal(z["n"], one.n) and np.array_equal(z["sum_x"], one.sum_x, equal_nan=True)
        assert np.array_equal(z["n_missing0"], one.n_missing[0])
        assert list(z["rsid"]) == list(one.rsid)
        for g in range(2):
            for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
                assert np.array_equal(z[f"{f}{g}"], one[f][g], equal_nan=True), (kernel, f, g)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_dense_and_packed_paths_on_a_second_device_in_one_process():
    """Kernel attributes (dynamic shared memory limits) are per device: after device 0 has run every path, device 1 must
    run them too inside the same process (the dense sweep needs 143-184 KB of dynamic shared memory)."""
    import hail_b200 as hb
    rng = np.random.default_rng(2)
    N, M = 3000, 600
    x = rng.integers(0, 3, size=(M, N)).astype(np.float64)
    x[rng.random(x.shape) < 0.03] = np.nan
    y, c = rng.normal(size=N), rng.normal(size=N)
    res = []
    for dev in (0, 1):
        d = hb.MatrixTable(hb.DenseDosage(x, device=dev), cols={"y": y, "c": c})
        hd = hb.linear_regression_rows(y=d.y, x=d.x, covariates=[1.0, d.c])
        g = hb.MatrixTable(hb.PackedGenotypes.from_dosage(np.where(np.isnan(x), -1, x).astype(np.int8), device=dev), cols={"y": y, "c": c})
        hp = [hb.linear_regression_rows(y=g.y, x=g.GT.n_alt_alleles(), covariates=[1.0, g.c], _kernel=k) for k in ("fp64", "tc4", "tc")]
        res.append((hd, hp))
    for f in ("beta", "standard_error", "p_value", "sum_x"):
        assert np.array_equal(res[0][0][f], res[1][0][f], equal_nan=True), f
        for a, b in zip(res[0][1], res[1][1]):
            assert np.array_equal(a[f], b[f], equal_nan=True), f
