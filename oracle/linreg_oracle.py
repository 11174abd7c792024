"""numpy restatement of Hail's per-variant linear regression (float64).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Every function cites the reference lines it follows.  ``LR`` =
``hail/hail/src/is/hail/methods/LinearRegression.scala``, ``RU`` =
``hail/hail/src/is/hail/stats/RegressionUtils.scala``, ``SG`` =
``hail/python/hail/methods/statgen.py``.

Conventions
-----------
* ``x``    : float64 ``[M, N]`` entry matrix (variants x samples); NaN = missing entry.
* ``ys``   : float64 ``[N, P]`` column phenotypes; NaN = missing.
* ``cov``  : float64 ``[N, K]`` column covariates; NaN = missing.
Outputs are dicts of numpy arrays with the reference's field names
(LR:26-34): ``n, sum_x, y_transpose_x, beta, standard_error, t_stat, p_value``.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import special


class OracleFatal(Exception):
    """Mirrors ``fatal(...)`` -> HailException -> Python FatalError (LR:55-58, RU:97-98, RU:113-114)."""


# --------------------------------------------------------------------------------------
# Student-t  (reference: jdistlib 0.4.5 ``T.cumulative`` = port of R nmath ``pt``; call
# sites LR:160, LR:344, stats/package.scala:373-374).  jdistlib is a third-party jar that
# is not under /root/reference, so this restates the published R algorithm:
#   val = pbeta(x^2/(n+x^2), 1/2, n/2, upper)   if n > x^2
#         pbeta(1/(1+x^2/n), n/2, 1/2, lower)   otherwise
#   P[T<=x] = val/2 for x<=0.   (The A&S 26.7.8 normal shortcut for n>4e5 is compiled out
#   of R since 2.6.0 -- `#ifdef R_version_le_260` -- so it is not restated.)
# --------------------------------------------------------------------------------------
def pt_lower(x, n):
    """P[T_n <= x] (``pT(x, n, lower_tail=True, log_p=False)``; functions.py:2626-2667)."""
    x = np.asarray(x, dtype=np.float64)
    n = np.asarray(n, dtype=np.float64)
    x2 = x * x
    with np.errstate(invalid="ignore", divide="ignore"):
        val = np.where(
            n > x2,
            special.betaincc(0.5, n / 2.0, x2 / (n + x2)),
            special.betainc(n / 2.0, 0.5, 1.0 / (1.0 + (x / n) * x)),
        )
    half = val / 2.0
    out = np.where(x <= 0.0, half, 1.0 - half)
    return np.where(np.isnan(x) | np.isnan(n), np.nan, out)


def two_sided_p(t, d):
    """``2 * T.cumulative(-|t|, d, lower=true, log=false)`` (LR:160, LR:344)."""
    t = np.asarray(t, dtype=np.float64)
    return 2.0 * pt_lower(-np.abs(t), float(d))


def _betacf(a, b, x, tol=3e-15, max_iter=200000):
    """Continued fraction for I_x(a,b) (modified Lentz).  Valid for x < (a+1)/(a+b+2)."""
    tiny = 1e-300
    qab, qap, qam = a + b, a + 1.0, a - 1.0
    c = 1.0
    d = 1.0 - qab * x / qap
    if abs(d) < tiny:
        d = tiny
    d = 1.0 / d
    h = d
    for m in range(1, max_iter + 1):
        m2 = 2 * m
        aa = m * (b - m) * x / ((qam + m2) * (a + m2))
        d = 1.0 + aa * d
        if abs(d) < tiny:
            d = tiny
        c = 1.0 + aa / c
        if abs(c) < tiny:
            c = tiny
        d = 1.0 / d
        h *= d * c
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2))
        d = 1.0 + aa * d
        if abs(d) < tiny:
            d = tiny
        c = 1.0 + aa / c
        if abs(c) < tiny:
            c = tiny
        d = 1.0 / d
        delta = d * c
        h *= delta
        if abs(delta - 1.0) < tol:
            return h
    raise RuntimeError("betacf failed to converge")


def log_two_sided_p(t, d):
    """Natural log of the two-sided p-value, usable far below 1e-308.

    p = I_{d/(d+t^2)}(d/2, 1/2) = x^a (1-x)^b / (a B(a,b)) * cf(a,b,x); evaluated in the
    log domain.  Scalar Python loop: small cases only.
    """
    t = abs(float(t))
    d = float(d)
    if math.isnan(t):
        return math.nan
    a, b = d / 2.0, 0.5
    t2d = t * t / d
    x = 1.0 / (1.0 + t2d)
    lbeta = math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)
    if d > 1e3:  # lgamma difference cancels badly for large a; use the Stirling series of log B(a, 1/2)
        lbeta = log_beta_half(a)
    if x < (a + 1.0) / (a + b + 2.0):
        log_front = -a * math.log1p(t2d) + b * math.log(t2d / (1.0 + t2d)) - math.log(a) - lbeta
        return log_front + math.log(_betacf(a, b, x))
    # near the centre: p = 1 - I_{1-x}(b, a)
    xc = t2d / (1.0 + t2d)
    if xc == 0.0:
        return 0.0
    log_front = b * math.log(xc) - a * math.log1p(t2d) - math.log(b) - lbeta
    lower = math.exp(log_front) * _betacf(b, a, xc)
    return math.log1p(-lower)


def log_beta_half(a):
    """log B(a, 1/2) = lgamma(a) + lgamma(1/2) - lgamma(a + 1/2), stable for large a.

    Uses lgamma(a) - lgamma(a+1/2) = -1/2 log a + 1/(8a) - 1/(192 a^3) + 1/(640 a^5) - ...
    (asymptotic series of the gamma ratio), exact to double precision for a >= 50.
    """
    if a < 50.0:
        return math.lgamma(a) + math.lgamma(0.5) - math.lgamma(a + 0.5)
    ia = 1.0 / a
    ia2 = ia * ia
    series = ia * (1.0 / 8.0 + ia2 * (-1.0 / 192.0 + ia2 * (1.0 / 640.0 + ia2 * (-17.0 / 14336.0))))
    return 0.5 * math.log(math.pi) - 0.5 * math.log(a) + series


# --------------------------------------------------------------------------------------
# RU:88-128  getPhenosCovCompleteSamples
# --------------------------------------------------------------------------------------
def complete_samples(ys, cov):
    """Keep samples where every phenotype and every covariate is defined (RU:100-110).

    Returns ``(y [n,P], cov [n,K], complete_col_idx int[n] ascending)``.
    """
    ys = np.asarray(ys, dtype=np.float64)
    cov = np.asarray(cov, dtype=np.float64)
    if ys.ndim != 2 or ys.shape[1] == 0:
        raise OracleFatal("No phenotypes present.")  # RU:97-98
    n_cols = ys.shape[0]
    cov = cov.reshape(n_cols, -1)
    keep = ~np.isnan(ys).any(axis=1) & ~np.isnan(cov).any(axis=1)
    idx = np.nonzero(keep)[0]
    if idx.size == 0:
        raise OracleFatal("No complete samples: each sample is missing its phenotype or some covariate")  # RU:113-114
    return ys[idx], cov[idx], idx


# --------------------------------------------------------------------------------------
# RU:16-58  setMeanImputedDoubles
# --------------------------------------------------------------------------------------
def mean_imputed_column(x_row, complete_col_idx):
    """One variant: select complete samples, sequential sum of defined entries, fill missing with the mean.

    ``sum`` runs in sample order exactly like RU:33-49 (python loop on purpose: it pins the
    summation order; only used for small cases -- ``mean_imputed_block`` is the vectorised form).
    """
    sel = x_row[complete_col_idx]
    n = sel.shape[0]
    out = np.empty(n, dtype=np.float64)
    s = 0.0
    missing = []
    for j in range(n):
        e = sel[j]
        if e != e:  # NaN == missing entry (RU:36,38 isElementDefined / isFieldDefined)
            missing.append(j)
        else:
            s += e
            out[j] = e
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.float64(s) / np.float64(n - len(missing))  # RU:52 (0/0 -> NaN for all-missing rows)
    out[missing] = mean
    return out


def mean_imputed_block(x_rows, complete_col_idx):
    """Vectorised RU:16-58 for a block of rows -> dense ``[n, B]`` column-major-like matrix X.

    For integer dosages the sequential float64 sum is exact, so this is bit-identical to
    ``mean_imputed_column``; for float dosages it is pinned by tests to within 1 ulp-ish.
    """
    sel = x_rows[:, complete_col_idx]  # [B, n]
    miss = np.isnan(sel)
    n = sel.shape[1]
    s = np.where(miss, 0.0, sel).sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = s / (n - miss.sum(axis=1))
    X = np.where(miss, mean[:, None], sel)
    return np.ascontiguousarray(X.T)  # [n, B]


# --------------------------------------------------------------------------------------
# LR:46-78 prologue and LR:134-160 block algebra
# --------------------------------------------------------------------------------------
def prologue(y, cov):
    """LR:47-78: n, K, d, Qt = justQ(qr.reduced(cov))^T, Qty, yyp."""
    n, k = y.shape[0], cov.shape[1]
    d = n - k - 1
    if d < 1:
        raise OracleFatal(
            f"{n} samples and {k + 1} {'covariate' if k == 1 else 'covariates'} (including x) implies {d} degrees of freedom."
        )  # LR:55-58
    if k > 0:
        q, _ = np.linalg.qr(cov, mode="reduced")  # LR:67
        Qt = np.ascontiguousarray(q.T)
    else:
        Qt = np.zeros((0, n))  # LR:69
    Qty = Qt @ y  # LR:71
    yyp = np.einsum("ij,ij->j", y, y) - np.einsum("ij,ij->j", Qty, Qty)  # LR:78
    return n, k, d, Qt, Qty, yyp


def block_algebra(X, y, Qt, Qty, yyp, d):
    """LR:134-160 on one block X [n, B].  Returns sum_x [B], ytx, b, se, t, p  each [P, B]."""
    AC = X.sum(axis=0)  # LR:136
    qtx = Qt @ X  # LR:139
    with np.errstate(invalid="ignore", divide="ignore"):
        xxp_rec = 1.0 / (np.einsum("ij,ij->j", X, X) - np.einsum("ij,ij->j", qtx, qtx))  # LR:141-142
        ytx = y.T @ X  # LR:143 (raw y)
        xyp = ytx - Qty.T @ qtx  # LR:146
        b = xyp * xxp_rec[None, :]  # LR:150-155
        se = np.sqrt((1.0 / d) * (np.outer(yyp, xxp_rec) - b * b))  # LR:157
        t = b / se  # LR:159
    p = two_sided_p(t, d)  # LR:160
    return AC, ytx, b, se, t, p


def linreg_group(x, ys, cov, block_size=16):
    """One group (= LinearRegressionRowsSingle.execute, LR:46-195).  Fields are ``[M]`` / ``[M, P]``."""
    x = np.asarray(x, dtype=np.float64)
    y, c, idx = complete_samples(ys, cov)
    n, k, d, Qt, Qty, yyp = prologue(y, c)
    M, P = x.shape[0], y.shape[1]
    out = {
        "n": np.full(M, n, dtype=np.int32),
        "sum_x": np.empty(M),
        "y_transpose_x": np.empty((M, P)),
        "beta": np.empty((M, P)),
        "standard_error": np.empty((M, P)),
        "t_stat": np.empty((M, P)),
        "p_value": np.empty((M, P)),
    }
    for r0 in range(0, M, block_size):  # trueGroupedIterator(rowBlockSize), LR:111
        r1 = min(M, r0 + block_size)
        X = mean_imputed_block(x[r0:r1], idx)
        AC, ytx, b, se, t, p = block_algebra(X, y, Qt, Qty, yyp, d)
        out["sum_x"][r0:r1] = AC
        out["y_transpose_x"][r0:r1] = ytx.T
        out["beta"][r0:r1] = b.T
        out["standard_error"][r0:r1] = se.T
        out["t_stat"][r0:r1] = t.T
        out["p_value"][r0:r1] = p.T
    out["_d"] = d
    return out


def linreg_group_weighted(x, ys, cov, w, block_size=16):
    """One group of `_linear_regression_rows_nd` with `weights` (statgen.py:530-545 kept samples: weight defined too;
    :557-581 sqrt-weight scaling of y and the covariates; :636-660 X = mean_impute(x) * sqrt(w), then the same algebra
    as LR:134-160 on the scaled quantities -- `sum_x` is the column sum of the SCALED X, statgen.py:646)."""
    x = np.asarray(x, dtype=np.float64)
    ys = np.asarray(ys, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    cov = np.asarray(cov, dtype=np.float64).reshape(ys.shape[0], -1)
    keep = ~np.isnan(ys).any(axis=1) & ~np.isnan(cov).any(axis=1) & ~np.isnan(w)
    idx = np.nonzero(keep)[0]
    if idx.size == 0:
        raise OracleFatal("No complete samples: each sample is missing its phenotype or some covariate")
    sw = np.sqrt(w[idx])
    y = ys[idx] * sw[:, None]
    c = cov[idx] * sw[:, None]
    n, k, d, Qt, Qty, yyp = prologue(y, c)
    M, P = x.shape[0], y.shape[1]
    out = {"n": np.full(M, n, dtype=np.int32), "sum_x": np.empty(M)}
    for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
        out[f] = np.empty((M, P))
    for r0 in range(0, M, block_size):
        r1 = min(M, r0 + block_size)
        X = mean_imputed_block(x[r0:r1], idx) * sw[:, None]          # statgen.py:637-644
        AC, ytx, b, se, t, p = block_algebra(X, y, Qt, Qty, yyp, d)
        out["sum_x"][r0:r1] = AC
        for f, v in (("y_transpose_x", ytx), ("beta", b), ("standard_error", se), ("t_stat", t), ("p_value", p)):
            out[f][r0:r1] = v.T
    out["_d"] = d
    return out


def linreg_chained(x, y_groups, cov, block_size=16):
    """LinearRegressionRowsChained.execute (LR:226-407): independent groups, outer array of length G.

    ``y_groups`` is a list of ``[N, P_g]`` arrays.  Returns a list of per-group dicts (the
    caller zips them into ``array<...>`` fields, LR:359-396).
    """
    return [linreg_group(x, yg, cov, block_size) for yg in y_groups]


# --------------------------------------------------------------------------------------
# Comparator used by Table._same (table.py:4384 -> utils/package.scala:234-252  D_==)
# --------------------------------------------------------------------------------------
MIN_NORMAL = 2.2250738585072014e-308


def d_eq(a, b, tolerance=1e-6):
    """a == b, both NaN, or |a-b| <= MIN_NORMAL + tol*max(|a|,|b|)  (elementwise)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        close = np.abs(a - b) <= MIN_NORMAL + tolerance * np.maximum(np.abs(a), np.abs(b))
    return (a == b) | (np.isnan(a) & np.isnan(b)) | close
