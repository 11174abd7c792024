"""`logistic_regression_rows` (wald / lrt / score / firth) -- host side of the B200 path.

Mirrors
  * Python API + validation   hail/python/hail/methods/statgen.py:731-1012 (defaults max_iterations=25, tolerance=1e-6;
                              ValueError for no covariates / empty y; same pass_through rules as the linear path)
  * driver prologue           hail/hail/src/is/hail/methods/LogisticRegression.scala:38-100: complete samples over ALL
                              phenotypes and covariates, 0/1 and non-constant checks, d >= 1, one null model fit per
                              phenotype (fatal when Newton does not converge), stats/LogisticRegressionModel.scala:279-370
  * output schema             key, pass_through, then the test's fields (stats/LogisticRegressionModel.scala:56-62, 111-116,
                              156-161, 212-215): wald `beta, standard_error, z_stat, p_value, fit`; lrt / firth `beta,
                              chi_sq_stat, p_value, fit`; score `chi_sq_stat, p_value`; `fit` = struct{n_iterations,
                              converged, exploded} (scalars for one y; for a list of y the reference nests them in
                              `logistic_regression: array<struct>` -- provided, next to flat [M, P] arrays)
The per-row loop (LogisticRegression.scala:115-157) runs in the CUDA library: the score test on the float64 sweep
(lrr_set_score_model / lrr_run_score), the Wald / LRT / Firth tests as per-variant Newton fits, one CTA per variant
(lrr_set_logit_model / lrr_run_logit, csrc/logit_kernel.cu).
"""
from __future__ import annotations

import ctypes
import logging
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .matrixtable import ChainedField, EntryExpression, ExpressionException, Table
from .statgen import FatalError, _column_values, _get_regression_row_fields, _plural, _warn_if_no_intercept

log = logging.getLogger("hail_b200")


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


_GRAM_SLICES = 16     # fixed, so that the summation order (and with it every bit of the null fit) does not depend on the host
_GRAM_THREADS = 4


def _weighted_gram(Ct, w, pool):
    """Ct diag(w) Ct' for covariate planes Ct [m, n]: the samples are cut into `_GRAM_SLICES` fixed slices whose partial Gram
    matrices are computed by a few host threads (numpy releases the GIL inside BLAS) and added in slice order."""
    m, n = Ct.shape
    edges = [n * i // _GRAM_SLICES for i in range(_GRAM_SLICES + 1)]

    def part(i):
        lo, hi = edges[i], edges[i + 1]
        a = Ct[:, lo:hi]
        return (a * w[lo:hi]) @ a.T

    parts = list(pool.map(part, range(_GRAM_SLICES))) if pool is not None else [part(i) for i in range(_GRAM_SLICES)]
    g = parts[0]
    for p_ in parts[1:]:
        g = g + p_
    return g


def _fit_null(C, y, max_iter, tol):
    """LogisticRegressionModel.fit from the intercept-only start (LogisticRegressionModel.scala:286-351).

    Works on covariate PLANES [m, n] (eta and the score are row-wise dot products; the Fisher matrix is a sliced, threaded
    Gram product): the n x m formulation with its n x m temporaries took 0.3 s at 10 covariates and 3.9 s at 63 for 400k
    samples -- longer than the device needs for a few hundred variants."""
    from concurrent.futures import ThreadPoolExecutor

    from .statgen import _blas_limits
    n, m = C.shape
    Ct = np.ascontiguousarray(C.T)
    b = np.zeros(m)
    avg = y.sum() / n
    b[0] = np.log(avg / (1.0 - avg))
    pool = ThreadPoolExecutor(_GRAM_THREADS) if n * m >= (1 << 20) else None
    try:
        with _blas_limits(limits=1):
            mu = _sigmoid(b @ Ct)
            score = Ct @ (y - mu)
            fisher = _weighted_gram(Ct, mu * (1.0 - mu), pool)
            it, converged, exploded = 0, False, False
            while not converged and not exploded and it < max_iter:
                it += 1
                try:
                    delta = np.linalg.solve(fisher, score)
                except np.linalg.LinAlgError:
                    exploded = True
                    break
                if np.isnan(delta[0]):
                    exploded = True
                elif np.max(np.abs(delta)) < tol:
                    converged = True
                else:
                    b = b + delta
                    mu = _sigmoid(b @ Ct)
                    score = Ct @ (y - mu)
                    fisher = _weighted_gram(Ct, mu * (1.0 - mu), pool)
    finally:
        if pool is not None:
            pool.shutdown()
    return b, mu, score, fisher, it, converged, exploded


def _fit_null_checked(C, y, max_iter, tol):
    fit = _fit_null(C, y, max_iter, tol)
    it, converged, exploded = fit[4], fit[5], fit[6]
    if not converged:   # LogisticRegression.scala:83-90
        raise FatalError("Failed to fit logistic regression null model (standard MLE with covariates only): " +
                         (f"exploded at Newton iteration {it}" if exploded else "Newton iteration failed to converge"))
    return fit


def logistic_regression_rows(test, y, x, covariates, pass_through=(), *, max_iterations=None, tolerance=None) -> Table:
    """For each row, test an input variable for association with a binary response using logistic regression
    (drop-in for `hl.logistic_regression_rows(test='score', ...)`, statgen.py:731)."""
    if test not in ("wald", "lrt", "score", "firth"):
        raise TypeError("logistic_regression_rows: parameter 'test': expected one of 'wald', 'lrt', 'score', 'firth', "
                        f"found {test!r}")
    if max_iterations is None:
        max_iterations = 25
    if tolerance is None:
        tolerance = 1e-6
    assert tolerance > 0.0
    if len(covariates) == 0:
        raise ValueError("logistic regression requires at least one covariate expression")
    if not isinstance(x, EntryExpression):
        raise ExpressionException("'logistic_regresion_rows/x': expected an entry-indexed expression "
                                  "(e.g. mt.GT.n_alt_alleles())")
    mt = x.source
    y_is_list = isinstance(y, list)
    if y_is_list and len(y) == 0:
        raise ValueError("'logistic_regression_rows': found no values for 'y'")
    ys = np.column_stack([_column_values(e, mt, "logistic_regression_rows/y") for e in (y if y_is_list else [y])])
    cov = np.column_stack([_column_values(e, mt, "logistic_regression_rows/covariates") for e in covariates])
    _warn_if_no_intercept("logistic_regression_rows", covariates)
    row_fields = _get_regression_row_fields(mt, pass_through, "logistic_regression_rows")

    # ---- LogisticRegression.scala:41-61 ----
    keep = ~np.isnan(ys).any(axis=1) & ~np.isnan(cov).any(axis=1)
    if not keep.any():
        raise FatalError("No complete samples: each sample is missing its phenotype or some covariate")
    yk, C = ys[keep], cov[keep]
    idx = np.ascontiguousarray(np.asarray(mt.col_index)[keep], dtype=np.int32)
    for col in range(yk.shape[1]):
        if not np.all((yk[:, col] == 0.0) | (yk[:, col] == 1.0)):
            raise FatalError(f"For logistic regression, y at index {col} must be bool or numeric with all present "
                             "values equal to 0 or 1")
        sy = yk[:, col].sum()
        if sy == 0.0 or sy == yk.shape[0]:
            raise FatalError(f"For logistic regression, y at index {col} must be non-constant")
    n, k = C.shape
    d = n - k - 1
    if d < 1:
        raise FatalError(f"{n} samples and {k + 1} {_plural(k, 'covariate')} (including x) implies {d} degrees of freedom.")
    log.info("logistic_regression_rows: running %s on %d samples for response variable y,\n"
             "    with input variable x, and %d additional %s...", test, n, k, _plural(k, "covariate"))

    from .genotypes import DenseDosage, HostBedGenotypes
    g = mt.genotypes
    dense = isinstance(g, DenseDosage)
    if dense != (x.kind == "dosage"):
        raise ExpressionException("'logistic_regression_rows/x': a dense dosage field needs a DenseDosage entry matrix")
    if isinstance(g, HostBedGenotypes):   # the score path sweeps a resident store
        g = g.to_device()
    dev = g.device
    ctx = _lib.context(dev.index)
    M, N, P = g.n_variants, g.n_samples, yk.shape[1]
    stream = torch.cuda.current_stream(dev).cuda_stream
    results = OrderedDict()
    with torch.cuda.device(dev):
        if test == "score":
            chi = np.empty((M, P))
            pv = np.empty((M, P))
            d_chi = torch.empty(M, dtype=torch.float64, device=dev)
            d_p = torch.empty(M, dtype=torch.float64, device=dev)
            out = _lib.ScoreOut(d_chi.data_ptr(), d_p.data_ptr(), None)
            for col in range(P):
                b, mu, score, fisher, it, converged, exploded = _fit_null_checked(C, yk[:, col], max_iterations, tolerance)
                w = mu * (1.0 - mu)
                wc = np.ascontiguousarray((C * w[:, None]).T)        # [K, n]
                resid = np.ascontiguousarray(yk[:, col] - mu)
                finv = np.ascontiguousarray(np.linalg.inv(fisher))
                ctx.check(ctx.lib.lrr_set_score_model(ctx.handle, N, n, k, idx.ctypes.data, wc.ctypes.data, resid.ctypes.data,
                                                      np.ascontiguousarray(w).ctypes.data, finv.ctypes.data,
                                                      np.ascontiguousarray(score).ctypes.data))
                if dense:
                    ctx.check(ctx.lib.lrr_run_score_dense(ctx.handle, g.data.data_ptr(), M, N, N, ctypes.byref(out), stream))
                else:
                    ctx.check(ctx.lib.lrr_run_score(ctx.handle, g.data.data_ptr(), g.flags_ptr(), M, g.stride, N,
                                                    ctypes.byref(out), stream))
                torch.cuda.synchronize(dev)
                chi[:, col] = d_chi.cpu().numpy()
                pv[:, col] = d_p.cpu().numpy()
            ctx.check(ctx.lib.lrr_clear_groups(ctx.handle))
            results["chi_sq_stat"] = chi
            results["p_value"] = pv
        else:
            names = {"wald": ("beta", "standard_error", "z_stat", "p_value"), "lrt": ("beta", "chi_sq_stat", "p_value"),
                     "firth": ("beta", "chi_sq_stat", "p_value")}[test]
            code = {"wald": 1, "lrt": 2, "firth": 3}[test]
            dev_out = {f: torch.empty(M, dtype=torch.float64, device=dev) for f in names}
            dev_out["n_iterations"] = torch.empty(M, dtype=torch.int32, device=dev)
            dev_out["converged"] = torch.empty(M, dtype=torch.uint8, device=dev)
            dev_out["exploded"] = torch.empty(M, dtype=torch.uint8, device=dev)
            out = _lib.LogitOut()
            for f, t in dev_out.items():
                setattr(out, f, t.data_ptr())
            host = {f: np.empty((M, P), dtype={"n_iterations": np.int32, "converged": bool, "exploded": bool}.get(f, np.float64))
                    for f in dev_out}
            Ct = np.ascontiguousarray(C.T)                              # [K, n]
            if k > 63:   # csrc/logit_kernel.cu: register form up to 19 covariates, tiled form (32 / 48 / 64 columns) up to 63
                raise ValueError(f"logistic_regression_rows: test={test!r} supports at most 63 covariates on the device "
                                 f"(found {k}); use test='score' for wider models")
            for col in range(P):
                b, mu, score, fisher, it, converged, exploded = _fit_null_checked(C, yk[:, col], max_iterations, tolerance)
                yc = np.ascontiguousarray(yk[:, col])
                with np.errstate(divide="ignore"):
                    loglik0 = float(np.sum(np.log(yc * mu + (1.0 - yc) * (1.0 - mu))))
                b_c, s_c, f_c = (np.ascontiguousarray(v, dtype=np.float64) for v in (b, score, fisher))
                ctx.check(ctx.lib.lrr_set_logit_model(ctx.handle, N, n, k, idx.ctypes.data, Ct.ctypes.data, yc.ctypes.data,
                                                      b_c.ctypes.data, s_c.ctypes.data, f_c.ctypes.data, loglik0))
                if dense:
                    ctx.check(ctx.lib.lrr_run_logit_dense(ctx.handle, g.data.data_ptr(), M, N, N, code, int(max_iterations),
                                                          float(tolerance), ctypes.byref(out), stream))
                else:
                    ctx.check(ctx.lib.lrr_run_logit(ctx.handle, g.data.data_ptr(), M, g.stride, N, code,
                                                    int(max_iterations), float(tolerance), ctypes.byref(out), stream))
                torch.cuda.synchronize(dev)
                for f, t in dev_out.items():
                    host[f][:, col] = t.cpu().numpy()
            for f in names:
                results[f] = host[f]
            results["fit"] = {f: host[f] for f in ("n_iterations", "converged", "exploded")}
    fields = OrderedDict()
    for kf in mt.row_key:
        fields[kf] = mt.row[kf]
    for kf, v in row_fields.items():
        fields[kf] = v
    for kf, v in results.items():
        if isinstance(v, dict):
            fields[kf] = {kk: (vv if y_is_list else vv[:, 0]) for kk, vv in v.items()}
        else:
            fields[kf] = v if y_is_list else v[:, 0]
    if y_is_list:
        # the reference's schema for a list of y: `logistic_regression: array<struct{...test fields...}>`, one struct per
        # phenotype (statgen.py:1003-1011); the flat [M, P] arrays above carry the same numbers
        fields["logistic_regression"] = ChainedField(
            OrderedDict((kf, ({kk: vv[:, p] for kk, vv in v.items()} if isinstance(v, dict) else v[:, p]))
                        for kf, v in results.items())
            for p in range(P))
    return Table(fields, key=mt.row_key, n_rows=mt.count_rows())
