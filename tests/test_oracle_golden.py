"""Pin the CPU oracle against the reference's own golden values.

Mirrors hail/python/test/hail/methods/test_statgen.py:223-284, 286-316, 366-457 and
hail/python/test/hail/expr/test_expr.py:3564-3568.
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import linreg_oracle as O
from tests.helpers import GOLDEN, gp_dosage, load_regression_linear, pl_dosage


def _check(res, expected, places=6):
    for pos, fields in expected.items():
        if not pos.isdigit():
            continue
        i = int(pos) - 1
        for f, v in fields.items():
            assert abs(res[f][i, 0] - v) < 0.5 * 10.0 ** (-places), (pos, f, res[f][i, 0], v)


def test_pt_known_answers():
    with open(os.path.join(GOLDEN, "pt_known_answers.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        p = float(O.pt_lower(c["x"], c["n"]))
        if not c["lower_tail"]:
            p = 1.0 - p
        if c["log_p"]:
            p = math.log(p)
        assert abs(p - c["value"]) <= 1e-14 * max(1.0, abs(c["value"])), c


def test_pt_against_scipy_stdtr():
    from scipy import special

    t = np.concatenate([np.linspace(-40, 0, 401), -np.logspace(-8, 1.5, 50)])
    for d in (1, 2, 5, 30, 994, 399989, 499989):
        a = O.two_sided_p(t, d)
        b = 2.0 * special.stdtr(d, -np.abs(t))
        ok = (b > 1e-300)
        # d=1, |t|=1e-8: exact value is 1-(2/pi)atan(t); scipy's stdtr is the one that is 3e-9 off there
        assert np.allclose(a[ok], b[ok], rtol=1e-8 if d == 1 else 1e-12, atol=0), d


def test_log_p_matches_p_where_representable_and_extends_below():
    for d in (2, 30, 994, 399989):
        for t in (0.0, 1e-3, 0.5, 1.7, 2.5, 5.0, 12.0, 30.0):
            p = float(O.two_sided_p(t, d))
            lp = O.log_two_sided_p(t, d)
            if p > 1e-300:
                assert abs(lp - math.log(p)) <= 1e-9 * max(1.0, abs(math.log(p))), (d, t, lp, math.log(p))
    # far tail: finite and monotone where the plain p-value has underflowed
    lp40 = O.log_two_sided_p(40.0, 399989)
    lp45 = O.log_two_sided_p(45.0, 399989)
    assert lp45 < lp40 < math.log(1e-300)
    # normal-limit sanity: log p(t=40, d=4e5) ~ log(2*Phi(-40)) within a few 1e-3 relative
    from scipy import special

    assert abs(lp40 - (math.log(2) + special.log_ndtr(-40.0))) < 0.02 * abs(lp40)


def test_linear_regression_without_intercept():  # TS:223-234
    x, y, cov, doc = load_regression_linear()
    res = O.linreg_group(x, y[:, None], np.empty((8, 0)))
    _check(res, doc["expected"]["no_intercept"])


def test_linear_regression_with_cov():  # TS:245-284
    x, y, cov, doc = load_regression_linear()
    covs = np.column_stack([np.ones(8), cov])
    res = O.linreg_group(x, y[:, None], covs)
    assert int(res["n"][0]) == 6 and res["_d"] == 2  # SURVEY 8c sample-alignment note
    exp = doc["expected"]["with_cov"]
    _check(res, exp)
    for v in exp["nan_se"]:
        assert np.isnan(res["standard_error"][v - 1, 0])
    for v in exp["nan_t_p"]:
        assert np.isnan(res["t_stat"][v - 1, 0]) and np.isnan(res["p_value"][v - 1, 0])


def test_linear_regression_pl():  # TS:286-316
    x, y, cov, doc = load_regression_linear()
    covs = np.column_stack([np.ones(8), cov])
    res = O.linreg_group(pl_dosage(doc), y[:, None], covs)
    _check(res, doc["expected"]["pl_dosage"])


def test_linear_regression_gp_dosage():  # TS:318-348 (places=4 on beta / standard_error, 6 on t / p)
    x, y, cov, doc = load_regression_linear()
    res = O.linreg_group(gp_dosage(doc), y[:, None], np.column_stack([np.ones(8), cov]))
    exp = doc["expected"]["gp_dosage"]
    for pos in ("1", "2", "3"):
        for f, v in exp[pos].items():
            tol = 5e-5 if f in ("beta", "standard_error") else 5e-7
            assert abs(res[f][int(pos) - 1, 0] - v) < tol, (pos, f)
    for v in exp["nan_se"]:
        assert np.isnan(res["standard_error"][v - 1, 0])


@pytest.mark.parametrize("kind", ["is_case", "quant"])
def test_linear_regression_with_import_fam(kind):  # TS:366-424
    x, y, cov, doc = load_regression_linear(fam_pheno=kind)
    covs = np.column_stack([np.ones(8), cov])
    res = O.linreg_group(x, y[:, None], covs)
    exp = {k: v for k, v in doc["expected"]["with_cov"].items() if k in ("1", "2")}
    if kind == "is_case":  # boolean phenotype = quant - 1: same beta / se / t / p
        pass
    _check(res, exp)
    for v in (6, 7, 8, 9, 10):
        assert np.isnan(res["standard_error"][v - 1, 0])


def test_multi_pheno_same_and_chained_equals_single():  # TS:426-457, TS:134-221
    x, y, cov, doc = load_regression_linear()
    covs = np.column_stack([np.ones(8), cov])
    single = O.linreg_group(x, y[:, None], covs)
    multi = O.linreg_group(x, np.column_stack([y, y]), covs)
    for f in ("beta", "standard_error", "t_stat", "p_value", "y_transpose_x"):
        assert O.d_eq(single[f][:, 0], multi[f][:, 0]).all()
        assert O.d_eq(multi[f][:, 0], multi[f][:, 1]).all()
    chained = O.linreg_chained(x, [y[:, None], y[:, None]], covs)
    for f in ("n", "sum_x", "beta", "standard_error", "t_stat", "p_value"):
        a, b = chained[0][f], chained[1][f]
        assert ((a == b) | (np.isnan(a) & np.isnan(b))).all()
    # differential missingness (TS:171-208): chained groups == separate filtered runs.
    # That reference test reads the pheno table WITHOUT missing='0' (TS:135), so X keeps pheno 0.0.
    x, y, cov, doc = load_regression_linear(pheno_missing_zero=False)
    y0 = np.where(cov[:, 1] >= 0, y, np.nan)
    y1 = np.where(cov[:, 1] <= 0, y, np.nan)
    c2 = np.column_stack([np.ones(8), cov[:, 0]])
    ch = O.linreg_chained(x, [y0[:, None], y1[:, None]], c2)
    for g, keep in enumerate([cov[:, 1] >= 0, cov[:, 1] <= 0]):
        sep = O.linreg_group(x[:, keep], y[keep][:, None], c2[keep])
        for f in ("n", "sum_x", "y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
            a, b = ch[g][f], sep[f]
            assert ((a == b) | (np.isnan(a) & np.isnan(b))).all(), (g, f)


def test_fatal_conditions():
    x, y, cov, doc = load_regression_linear()
    with pytest.raises(O.OracleFatal, match="degrees of freedom"):  # LR:55-58
        O.linreg_group(x, y[:, None], np.column_stack([np.ones(8), cov, cov[:, 0] ** 2, cov[:, 1] ** 2, cov[:, 0] * cov[:, 1]]))
    with pytest.raises(O.OracleFatal, match="No complete samples"):  # RU:113-114
        O.linreg_group(x, np.full((8, 1), np.nan), np.ones((8, 1)))
    with pytest.raises(O.OracleFatal, match="No phenotypes"):  # RU:97-98
        O.linreg_group(x, np.empty((8, 0)), np.ones((8, 1)))


def test_sequential_and_vectorised_imputation_agree():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 3, size=(40, 57)).astype(np.float64)
    x[rng.random(x.shape) < 0.2] = np.nan
    x[3] = np.nan  # all-missing row -> NaN mean (RU:52)
    idx = np.sort(rng.choice(57, 41, replace=False))
    X = O.mean_imputed_block(x, idx)
    for r in range(40):
        col = O.mean_imputed_column(x[r], idx)
        assert np.array_equal(col, X[:, r], equal_nan=True)


def test_weighted_oracle_equals_preweighted_ols():
    """test_statgen.py:553-593: regression with weights w == OLS on y, x, covariates all scaled by sqrt(w) (no missing
    calls, as the reference test arranges with coalesce)."""
    rng = np.random.default_rng(3)
    N, M = 60, 25
    x = rng.integers(0, 3, size=(M, N)).astype(np.float64)
    y = rng.normal(size=(N, 2))
    cov = np.column_stack([np.ones(N), rng.normal(size=(N, 2))])
    w = rng.uniform(0.5, 3.0, size=N)
    got = O.linreg_group_weighted(x, y, cov, w)
    sw = np.sqrt(w)
    want = O.linreg_group(x * sw[None, :], y * sw[:, None], cov * sw[:, None])
    for f in ("sum_x", "y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
        assert np.allclose(got[f], want[f], rtol=1e-10, atol=0), f
    # missing weights drop the sample (test_statgen.py:610-660)
    w2 = w.copy()
    w2[[3, 17]] = np.nan
    keep = ~np.isnan(w2)
    a = O.linreg_group_weighted(x, y, cov, w2)
    b = O.linreg_group_weighted(x[:, keep], y[keep], cov[keep], w2[keep])
    assert (a["n"] == N - 2).all()
    for f in ("beta", "p_value"):
        assert np.allclose(a[f], b[f], rtol=1e-12, atol=0), f
