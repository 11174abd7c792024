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
