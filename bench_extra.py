"""Secondary measurements: the neighbours of the hot path on one B200, one JSON line each.

    python bench_extra.py [--which dense,logistic,pca] [--samples N]

Not part of the driver's bench contract (that is bench.py: `linear_regression_rows` on C2); same conventions -- inputs
resident in HBM and larger than L2, CUDA events on the launching stream, 3 warm-up calls -- and, for the HBM-bound dense
sweep, the same `roofline` object (algorithmic bytes = 8 B per entry + result rows).
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 7700.0, "nominal fallback"


def covariates(n, k, seed=0):
    rng = np.random.default_rng(seed)
    return np.column_stack([np.ones(n)] + [rng.normal(size=n) for _ in range(k - 1)]), rng


def bench_dense(N, M, K, missing, compact=False):
    import hail_b200 as hb
    from hail_b200 import _lib, statgen
    from hail_b200.statgen import GroupBasis
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    cov, rng = covariates(N, K)
    y = rng.normal(size=(N, 1))
    if compact:   # uniform dosages as uint16 entries (value = q * scale, 0xFFFF = missing), drawn on the device row block by row block
        scale = 2.0 / 65534.0
        ld = (N + 7) // 8 * 8
        xq = torch.zeros((M, ld), dtype=torch.int16, device=dev)
        for lo in range(0, M, 1024):
            hi = min(M, lo + 1024)
            q = torch.randint(0, 65535, (hi - lo, N), device=dev, generator=g, dtype=torch.int32)
            q[torch.rand((hi - lo, N), device=dev, generator=g) < missing] = 65535
            xq[lo:hi, :N] = torch.where(q >= 32768, q - 65536, q).to(torch.int16)    # the bits of the uint16 value
        del q
    else:
        x = torch.rand((M, N), device=dev, dtype=torch.float64, generator=g) * 2
        x[torch.rand((M, N), device=dev, generator=g) < missing] = float("nan")
        dd = hb.DenseDosage(x, 0)
    ctx = _lib.context(0)
    statgen._push_groups(ctx, N, [GroupBasis(y, cov, np.arange(N), None)])
    o = {"n": torch.empty(M, dtype=torch.int32, device=dev), "n_missing": torch.empty(M, dtype=torch.int32, device=dev),
         "sum_x": torch.empty(M, dtype=torch.float64, device=dev)}
    for f in statgen.STAT_FIELDS:
        o[f] = torch.empty((M, 1), dtype=torch.float64, device=dev)
    arr = (_lib.GroupOut * 1)()
    for k, v in o.items():
        setattr(arr[0], k, v.data_ptr())
    arr[0].log10_p = None
    stream = torch.cuda.current_stream(dev).cuda_stream

    def run():
        if compact:
            ctx.check(ctx.lib.lrr_run_dense_u16(ctx.handle, xq.data_ptr(), M, ld, N, scale, arr, 1, stream))
        else:
            ctx.check(ctx.lib.lrr_run_dense(ctx.handle, dd.data.data_ptr(), M, N, N, arr, 1, stream))

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_alg = M * N * (2 if compact else 8) + M * (4 + 4 + 8 + 5 * 8)
    peak, src = hbm_peak()
    ach = bytes_alg / ms / 1e6
    return {"metric": "entries/sec (variants x samples) for linear_regression_rows on a dense " + ("uint16 (compact)" if compact else "float64") + " x",
            "value": M * N / ms * 1e3,
            "unit": "entries/s", "n_gpus": 1, "steps": reps, "warmup": 3, "ms_per_step": ms, "higher_is_better": True,
            "dtype": "f64 arithmetic" + (", u16 storage" if compact else ""), "data": "synthetic (uniform dosages in [0, 2], generated in HBM)",
            "config": {"workload": f"dense x ({'uint16' if compact else 'float64'} entries): {N} samples x {M} variants, P=1, K={K}", "missing_rate": missing,
                       "l2": "inputs larger than L2 (%.1f GB)" % (M * N * (2 if compact else 8) / 1e9)},
            "gpu_launches": int(ctx.launch_count - l0),
            "roofline": {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                         "traffic": None, "kernel": "dense sweep + imputation + statistics (whole call)",
                         "algorithmic_bytes_per_launch": bytes_alg, "peak_source": src}}


def bench_logistic(N, M, K, test):
    import hail_b200 as hb
    mt = hb.balding_nichols_model(3, N, M, missing_rate=0.01, seed=5)
    cov, rng = covariates(N, K)
    y = (rng.random(N) < 1 / (1 + np.exp(-(0.3 * cov[:, 1] - 0.2)))).astype(np.float64)
    mt = mt.annotate_cols(y=y, **{f"c{k}": cov[:, k] for k in range(1, K)})
    covs = [1.0] + [mt[f"c{k}"] for k in range(1, K)]
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ht = hb.logistic_regression_rows(test, mt.y, mt.GT.n_alt_alleles(), covs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    it = float(ht.fit["n_iterations"].mean()) if test != "score" else 0.0
    # CPU baseline: the oracle's numpy restatement of the reference's per-row loop (LogisticRegression.scala:115-157) on a
    # bounded sample of the same rows, host BLAS threads as configured ("port": the JVM reference cannot run here)
    from oracle import logreg_oracle as LO
    Ms = (8 if K <= 12 else 2) if test != "score" else 32
    dos = mt.genotypes.rows(0, Ms).to_dosage().astype(np.float64)
    dos[dos < 0] = np.nan
    t0 = time.perf_counter()
    LO.logreg_score(dos, y, cov) if test == "score" else LO.logreg_rows(test, dos, y, cov)
    dt_cpu = time.perf_counter() - t0
    return {"metric": f"genotypes/sec for logistic_regression_rows(test='{test}') through the public call (host null fit included)",
            "value": N * M / dt, "unit": "genotypes/s", "variants_per_s": M / dt, "n_gpus": 1, "seconds": dt, "higher_is_better": True,
            "cpu_baseline": {"value": N * Ms / dt_cpu, "unit": "genotypes/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{Ms} variants x {N} samples, {dt_cpu:.1f} s, oracle/logreg_oracle.py (numpy, null fit included)"},
            "dtype": "f64", "data": "synthetic (seeded Balding-Nichols style, 1 % missing)", "mean_newton_iterations": it,
            "config": {"workload": f"logistic {test}: {N} samples x {M} variants, K={K}"}}


def bench_pca(N, M, k):
    import hail_b200 as hb
    mt = hb.balding_nichols_model(k + 1, N, M, missing_rate=0.01, seed=5)
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev, scores, _ = hb.hwe_normalized_pca(mt.GT, k=k)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return {"metric": "seconds for hwe_normalized_pca through the public call", "value": dt, "unit": "s", "n_gpus": 1,
            "higher_is_better": False, "dtype": "f64 (sweep: exact-integer tensor-core digits)", "iterations": int(scores.n_iterations),
            "converged": bool(scores.converged), "passes_over_genotypes": 2 * int(scores.n_iterations) + 1,
            "data": "synthetic (seeded Balding-Nichols style, k + 1 populations, 1 % missing)",
            "eigenvalues": [round(float(v), 4) for v in ev],
            "config": {"workload": f"PCA: {N} samples x {M} variants, k={k}"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="dense,logistic,pca")
    ap.add_argument("--samples", type=int, default=400000)
    a = ap.parse_args()
    N = a.samples
    which = a.which.split(",")
    if "dense" in which:
        for miss in (0.0, 0.01):
            print(json.dumps(bench_dense(N, 4736, 10, miss)), flush=True)
        for miss in (0.0, 0.01):
            print(json.dumps(bench_dense(N, 4736 * 4, 10, miss, compact=True)), flush=True)
    if "logistic" in which:
        for test in ("score", "wald", "lrt", "firth"):
            print(json.dumps(bench_logistic(N, 2048, 10, test)), flush=True)
    if "logistic_wide" in which:   # wide models: the register form's last size (19 covariates) and the tiled form (32 / 48 / 64 columns)
        for K in (19, 24, 40, 63):
            print(json.dumps(bench_logistic(N, 592, K, "wald")), flush=True)
    if "logistic_mid" in which:   # 12 - 19 covariates: register form (2 / 4 slices) against the tiled form (tuning builds: LRR_LOGIT_TILED_FROM)
        for K in (10, 12, 15, 19):
            print(json.dumps(bench_logistic(N, 592, K, "wald")), flush=True)
    if "pca" in which:
        print(json.dumps(bench_pca(N, 100000, 5)), flush=True)


if __name__ == "__main__":
    main()
