"""The logistic score-test oracle against the reference's R-derived golden values (test_statgen.py:987-1021)."""
import json
import os

import numpy as np
import pytest

from oracle import logreg_oracle as L
from oracle.linreg_oracle import OracleFatal

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_regression_logistic():
    doc = json.load(open(os.path.join(GOLDEN, "regression_logistic.json")))
    s = doc["samples"]
    x = np.array([[np.nan if v is None else float(v) for v in row] for row in doc["gt_n_alt_alleles"]])
    y = np.array([np.nan if doc["pheno_table"].get(k) is None else float(doc["pheno_table"][k]) for k in s])
    cov = np.array([[1.0] + doc["cov_table"].get(k, [np.nan, np.nan]) for k in s])
    return doc, x, y, cov


def test_score_test_matches_r():
    doc, x, y, cov = load_regression_logistic()
    out = L.logreg_score(x, y, cov)
    assert out["n"] == 10                                  # A..J minus W, X (no phenotype / covariates)
    for pos, want in doc["expected_score"].items():
        if pos == "constant":
            continue
        i = int(pos) - 1
        assert abs(out["chi_sq_stat"][i] - want["chi_sq_stat"]) < 5e-7, pos
        assert abs(out["p_value"][i] - want["p_value"]) < 5e-7, pos
    for pos in doc["expected_score"]["constant"]:
        c = out["chi_sq_stat"][pos - 1]
        assert np.isnan(c) or c < 1e-6


def test_null_fit_and_fatal_conditions():
    _, x, y, cov = load_regression_logistic()
    with pytest.raises(OracleFatal, match="must be non-constant"):
        L.logreg_score(x, np.where(np.isnan(y), np.nan, 1.0), cov)
    with pytest.raises(OracleFatal, match="equal to 0 or 1"):
        L.logreg_score(x, np.where(np.isnan(y), np.nan, y * 2.0), cov)
    with pytest.raises(OracleFatal, match="degrees of freedom"):
        L.logreg_score(x, y, np.column_stack([cov] + [np.random.default_rng(i).normal(size=cov.shape[0]) for i in range(7)]))
    # separable covariate -> the null model explodes / does not converge (LogisticRegression.scala:83-90)
    sep = np.column_stack([cov[:, 0], np.where(np.nan_to_num(y) > 0, 5.0, -5.0)])
    with pytest.raises(OracleFatal, match="Failed to fit logistic regression null model"):
        L.logreg_score(x, y, sep)
    assert np.allclose(L.chi_sq_tail_1(np.array([0.0, 3.841458820694124])), [1.0, 0.05], atol=1e-12)


# ---- wald / lrt / firth (per-variant Newton fits) ------------------------------------------------------------
def test_wald_and_lrt_match_r():   # test_statgen.py:719-756, 940-985
    doc, x, y, cov = load_regression_logistic()
    for test, key in (("wald", "expected_wald"), ("lrt", "expected_lrt")):
        out = L.logreg_rows(test, x, y, cov)
        exp = doc[key]
        for pos in ("1", "2"):
            for f, v in exp[pos].items():
                assert abs(out[f][int(pos) - 1] - v) < 5e-7, (test, pos, f)
        for pos in exp["not_converged"]:
            assert not out["converged"][pos - 1]                # separable
        for pos in exp["constant"]:
            i = pos - 1
            assert (not out["converged"][i]) or np.isnan(out["p_value"][i]) or abs(out["p_value"][i] - 1) < 1e-4


def load_epacts():
    z = np.load(os.path.join(GOLDEN, "logistic_epacts.npz"))
    x = z["gt"].astype(np.float64)
    x[x < 0] = np.nan
    cov = np.column_stack([np.ones(x.shape[1]), z["is_female"], z["pc1"], z["pc2"]])
    return z, x, z["is_case"], cov


def test_epacts_all_four_tests():  # test_statgen.py:1722-1862
    z, x, y, cov = load_epacts()
    w = L.logreg_rows("wald", x, y, cov)
    for i in range(5):
        for j, f in enumerate(("beta", "standard_error", "z_stat", "p_value")):
            assert w[f][i] == pytest.approx(z["wald"][i, j], rel=z["wald_rel"][i, j]), (i, f)
    lrt = L.logreg_rows("lrt", x, y, cov)
    assert lrt["p_value"] == pytest.approx(z["lrt_p"], rel=1e-4)
    sc = L.logreg_score(x, y, cov)
    assert sc["chi_sq_stat"] == pytest.approx(z["score"][:, 0], rel=1e-5)
    assert sc["p_value"] == pytest.approx(z["score"][:, 1], rel=1e-5)
    fi = L.logreg_rows("firth", x, y, cov)
    assert fi["beta"] == pytest.approx(z["firth"][:, 0], rel=1e-4)
    assert fi["p_value"] == pytest.approx(z["firth"][:, 1], rel=1e-4)
    assert fi["converged"].all() and w["converged"].all()


@pytest.mark.parametrize("which", ["pl", "gp"])
def test_wald_on_dosages(which):   # test_statgen.py:851-938
    from tests.helpers import gp_dosage, pl_dosage
    doc, x, y, cov = load_regression_logistic()
    dos = pl_dosage(doc) if which == "pl" else gp_dosage(doc)
    out = L.logreg_rows("wald", dos, y, cov)
    exp = doc["expected_wald_dosage"]
    tol = 5e-7 if which == "pl" else 5e-5
    for pos in ("1", "2"):
        for f, v in exp[pos].items():
            assert abs(out[f][int(pos) - 1] - v) < tol, (which, pos, f, out[f][int(pos) - 1])
    assert not out["converged"][2]
    for pos in exp["constant"]:
        i = pos - 1
        assert (not out["converged"][i]) or np.isnan(out["p_value"][i]) or abs(out["p_value"][i] - 1) < 1e-4
