"""CPU restatement of the reference's per-variant logistic regression SCORE test.  TEST INFRASTRUCTURE ONLY:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Follows
  hail/hail/src/is/hail/methods/LogisticRegression.scala:38-100   complete samples, 0/1 and non-constant checks, d >= 1,
                                                                  null model fit (fatal when it does not converge)
  hail/hail/src/is/hail/stats/LogisticRegressionModel.scala:279-370   Newton iterations from the intercept-only start
  hail/hail/src/is/hail/stats/LogisticRegressionModel.scala:211-264   LogisticScoreTest: score and Fisher information
                                                                  of [covariates | x] at the null fit, chi2 = s' F^-1 s,
                                                                  p = pchisqtail(chi2, 1)
  hail/hail/src/is/hail/stats/RegressionUtils.scala:16-58         x is mean-imputed over the complete samples
Pinned by the R-derived values of hail/python/test/hail/methods/test_statgen.py:987-1021
(tests/golden/regression_logistic.json, tests/test_oracle_logistic.py).
"""
import math

import numpy as np

from .linreg_oracle import OracleFatal, complete_samples, mean_imputed_block


def sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def fit_null(C, y, max_iter=25, tol=1e-6):
    """LogisticRegressionModel.fit with no null fit (LogisticRegressionModel.scala:294-370).
    Returns dict(b, score, fisher, mu, n_iter, converged, exploded)."""
    n, m = C.shape
    b = np.zeros(m)
    avg = y.sum() / n
    b[0] = math.log(avg / (1.0 - avg))                  # bInterceptOnly (column 0 is taken to be the intercept)
    mu = sigmoid(C @ b)
    score = C.T @ (y - mu)
    fisher = C.T @ (C * (mu * (1.0 - mu))[:, None])
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
            mu = sigmoid(C @ b)
            score = C.T @ (y - mu)
            fisher = C.T @ (C * (mu * (1.0 - mu))[:, None])
    return {"b": b, "score": score, "fisher": fisher, "mu": mu, "n_iter": it, "converged": converged, "exploded": exploded}


def chi_sq_tail_1(chi2):
    """pchisqtail(chi2, 1) = erfc(sqrt(chi2 / 2))."""
    chi2 = np.asarray(chi2, dtype=np.float64)
    out = np.full(chi2.shape, np.nan)
    ok = chi2 >= 0
    out[ok] = [math.erfc(math.sqrt(v / 2.0)) for v in chi2[ok]]
    out[chi2 < 0] = 1.0          # a roundoff-negative statistic: the upper tail is the whole distribution (R: pchisq)
    return out


def logreg_score(x, y, cov, max_iter=25, tol=1e-6):
    """One phenotype.  x [M, N] dosages (NaN = missing), y [N] 0/1 (NaN = missing), cov [N, K] (column 0 the intercept).
    Returns dict(n, chi_sq_stat [M], p_value [M]); NaN where the reference yields missing (singular Fisher matrix)."""
    x = np.asarray(x, dtype=np.float64)
    yy, C, idx = complete_samples(np.asarray(y, dtype=np.float64).reshape(-1, 1), cov)
    yv = yy[:, 0]
    if not np.all((yv == 0.0) | (yv == 1.0)):
        raise OracleFatal("For logistic regression, y at index 0 must be bool or numeric with all present values equal to 0 or 1")
    if yv.sum() == 0.0 or yv.sum() == yv.size:
        raise OracleFatal("For logistic regression, y at index 0 must be non-constant")
    n, k = C.shape
    d = n - k - 1
    if d < 1:
        raise OracleFatal(f"{n} samples and {k + 1} {'covariate' if k == 1 else 'covariates'} (including x) implies {d} degrees of freedom.")
    nf = fit_null(C, yv, max_iter, tol)
    if not nf["converged"]:
        raise OracleFatal("Failed to fit logistic regression null model (standard MLE with covariates only): " +
                          (f"exploded at Newton iteration {nf['n_iter']}" if nf["exploded"]
                           else "Newton iteration failed to converge"))
    mu = nf["mu"]
    w = mu * (1.0 - mu)
    M = x.shape[0]
    chi2 = np.full(M, np.nan)
    X = mean_imputed_block(x, idx)                       # [n, M]
    for v in range(M):
        xv = X[:, v]
        score = np.concatenate([nf["score"], [xv @ (yv - mu)]])
        fisher = np.empty((k + 1, k + 1))
        fisher[:k, :k] = nf["fisher"]
        fisher[:k, k] = C.T @ (xv * w)
        fisher[k, :k] = fisher[:k, k]
        fisher[k, k] = xv @ (xv * w)
        try:
            with np.errstate(all="ignore"):
                sol = np.linalg.solve(fisher, score)
            chi2[v] = score @ sol
        except np.linalg.LinAlgError:
            pass                                           # MatrixSingularException -> missing (scala :256-259)
    return {"n": n, "chi_sq_stat": chi2, "p_value": chi_sq_tail_1(chi2)}


# --------------------------------------------------------------------------------------------------------------
# Wald / likelihood-ratio / Firth tests: per-variant Newton fits (LogisticRegressionModel.scala:55-207, 294-408)
# --------------------------------------------------------------------------------------------------------------
def _loglik(y, mu):
    with np.errstate(divide="ignore"):
        return float(np.sum(np.log(y * mu + (1.0 - y) * (1.0 - mu))))     # scala :365


def fit_with_null(X, y, nf, max_iter=25, tol=1e-6):
    """LogisticRegressionModel.fit(Some(nullFit)) (scala :294-370): the full model [covariates | x] started from the
    null fit; score / Fisher blocks of the covariates are taken over from the null fit for the first step (:311-325)."""
    n, m = X.shape
    m0 = nf["b"].size
    b = np.zeros(m)
    b[:m0] = nf["b"]
    mu = sigmoid(X @ b)
    w = mu * (1.0 - mu)
    score = np.empty(m)
    score[:m0] = nf["score"]
    score[m0:] = X[:, m0:].T @ (y - mu)
    fisher = np.empty((m, m))
    fisher[:m0, :m0] = nf["fisher"]
    fisher[:m0, m0:] = X[:, :m0].T @ (X[:, m0:] * w[:, None])
    fisher[m0:, :m0] = fisher[:m0, m0:].T
    fisher[m0:, m0:] = X[:, m0:].T @ (X[:, m0:] * w[:, None])
    it, converged, exploded = 0, False, False
    while not converged and not exploded and it < max_iter:
        it += 1
        try:
            with np.errstate(all="ignore"):
                delta = np.linalg.solve(fisher, score)
        except np.linalg.LinAlgError:                       # MatrixSingularException (:360-361)
            exploded = True
            break
        if np.isnan(delta[0]):
            exploded = True
        elif np.max(np.abs(delta)) < tol:
            converged = True
        else:
            b = b + delta
            with np.errstate(over="ignore"):
                mu = sigmoid(X @ b)
            score = X.T @ (y - mu)
            fisher = X.T @ (X * (mu * (1.0 - mu))[:, None])
    return {"b": b, "score": score, "fisher": fisher, "mu": mu, "loglik": _loglik(y, mu), "n_iter": it,
            "converged": converged, "exploded": exploded}


def fit_firth(X, y, b0, max_iter=25, tol=1e-6):
    """LogisticRegressionModel.fitFirth (scala :372-408): b has len(b0) free coefficients, the hat diagonal comes from
    the QR of the FULL sqrt(W)-scaled design; converged only after the first iteration (:393)."""
    n, m = X.shape
    b = np.array(b0, dtype=np.float64)
    m0 = b.size
    loglik, it, converged, exploded = 0.0, 0, False, False
    while not converged and not exploded and it < max_iter:
        it += 1
        with np.errstate(all="ignore"):
            mu = sigmoid(X[:, :m0] @ b)
            sqrt_w = np.sqrt(mu * (1.0 - mu))
            try:
                Q, R = np.linalg.qr(X * sqrt_w[:, None])
            except np.linalg.LinAlgError:
                exploded = True
                break
            h = np.sum(Q * Q, axis=1)
            rhs = Q[:, :m0].T @ (((y - mu) + h * (0.5 - mu)) / sqrt_w)
            try:
                delta = np.linalg.solve(R[:m0, :m0], rhs)      # TriSolve: an exactly zero pivot is singular
                if np.any(np.diag(R[:m0, :m0]) == 0.0):
                    raise np.linalg.LinAlgError
            except np.linalg.LinAlgError:
                exploded = True
                break
        if np.isnan(delta[0]):
            exploded = True
        elif np.max(np.abs(delta)) < tol and it > 1:
            converged = True
            with np.errstate(divide="ignore"):
                loglik = _loglik(y, mu) + float(np.sum(np.log(np.abs(np.diag(R)))))
        else:
            b = b + delta
    return {"b": b, "loglik": loglik, "n_iter": it, "converged": converged, "exploded": exploded}


def _prepare(x, y, cov, max_iter, tol):
    x = np.asarray(x, dtype=np.float64)
    yy, C, idx = complete_samples(np.asarray(y, dtype=np.float64).reshape(-1, 1), cov)
    yv = yy[:, 0]
    if not np.all((yv == 0.0) | (yv == 1.0)):
        raise OracleFatal("For logistic regression, y at index 0 must be bool or numeric with all present values equal to 0 or 1")
    if yv.sum() == 0.0 or yv.sum() == yv.size:
        raise OracleFatal("For logistic regression, y at index 0 must be non-constant")
    n, k = C.shape
    d = n - k - 1
    if d < 1:
        raise OracleFatal(f"{n} samples and {k + 1} {'covariate' if k == 1 else 'covariates'} (including x) implies {d} degrees of freedom.")
    nf = fit_null(C, yv, max_iter, tol)
    if not nf["converged"]:
        raise OracleFatal("Failed to fit logistic regression null model (standard MLE with covariates only): " +
                          (f"exploded at Newton iteration {nf['n_iter']}" if nf["exploded"]
                           else "Newton iteration failed to converge"))
    nf["loglik"] = _loglik(yv, nf["mu"])
    return x, yv, C, idx, nf


def logreg_rows(test, x, y, cov, max_iter=25, tol=1e-6):
    """One phenotype, test in {'wald', 'lrt', 'firth'} (LogisticRegression.scala:115-157 with WaldTest :55-93,
    LikelihoodRatioTest :110-145, LogisticFirthTest :155-199).  Fields are [M]; NaN where the reference leaves the
    statistics missing (fit not converged / singular), `fit` fields always present."""
    x, yv, C, idx, nf = _prepare(x, y, cov, max_iter, tol)
    M = x.shape[0]
    k = C.shape[1]
    X_imp = mean_imputed_block(x, idx)                  # [n, M]
    names = {"wald": ("beta", "standard_error", "z_stat", "p_value"), "lrt": ("beta", "chi_sq_stat", "p_value"),
             "firth": ("beta", "chi_sq_stat", "p_value")}[test]
    out = {f: np.full(M, np.nan) for f in names}
    out["n_iterations"] = np.zeros(M, dtype=np.int32)
    out["converged"] = np.zeros(M, dtype=bool)
    out["exploded"] = np.zeros(M, dtype=bool)
    out["n"] = C.shape[0]
    for v in range(M):
        X = np.column_stack([C, X_imp[:, v]])
        if test == "firth":
            f0 = fit_firth(X, yv, nf["b"], max_iter, tol)
            fit = f0
            if f0["converged"]:
                f1 = fit_firth(X, yv, np.concatenate([f0["b"], [0.0]]), max_iter, tol)
                fit = f1
                if f1["converged"]:
                    chi2 = 2.0 * (f1["loglik"] - f0["loglik"])
                    out["beta"][v] = f1["b"][-1]
                    out["chi_sq_stat"][v] = chi2
                    out["p_value"][v] = chi_sq_tail_1(np.array([chi2]))[0]
        else:
            fit = fit_with_null(X, yv, nf, max_iter, tol)
            if fit["converged"]:
                if test == "wald":
                    try:
                        with np.errstate(all="ignore"):
                            se = np.sqrt(np.diag(np.linalg.inv(fit["fisher"])))
                        z = fit["b"] / se
                        out["beta"][v] = fit["b"][-1]
                        out["standard_error"][v] = se[-1]
                        out["z_stat"][v] = z[-1]
                        out["p_value"][v] = math.erfc(abs(z[-1]) / math.sqrt(2.0)) if not np.isnan(z[-1]) else np.nan
                    except np.linalg.LinAlgError:
                        pass
                else:
                    chi2 = 2.0 * (fit["loglik"] - nf["loglik"])
                    out["beta"][v] = fit["b"][-1]
                    out["chi_sq_stat"][v] = chi2
                    out["p_value"][v] = chi_sq_tail_1(np.array([chi2]))[0]
        out["n_iterations"][v] = fit["n_iter"]
        out["converged"][v] = fit["converged"]
        out["exploded"][v] = fit["exploded"]
    return out
