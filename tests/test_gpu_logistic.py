"""GPU parity of `logistic_regression_rows(test='wald' | 'lrt' | 'firth')` (lrr_set_logit_model / lrr_run_logit):
the reference's R / EPACTS golden values (test_statgen.py:719-756, 940-985, 1722-1862) and the CPU oracle."""
import os

import numpy as np
import pytest

from oracle import bed as obed
from oracle import logreg_oracle as L
from tests.helpers import GOLDEN
from tests.test_oracle_logistic import load_epacts, load_regression_logistic

pytestmark = pytest.mark.gpu


def _hb():
    import hail_b200 as hb
    return hb


def _mt(x, **cols):
    hb = _hb()
    gt = hb.PackedGenotypes.from_dosage(np.where(np.isnan(x), -1, x).astype(np.int8))
    return hb.MatrixTable(gt, rows={"idx": np.arange(x.shape[0])}, cols=cols)


def _close(a, b, rel, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    ok = both_nan | (np.abs(a - b) <= 1e-12 + rel * np.maximum(np.abs(a), np.abs(b)))
    assert ok.all(), f"{what}: {np.count_nonzero(~ok)} mismatches, first {np.argwhere(~ok)[:3].tolist()}, " \
                     f"{a[~ok][:3]} vs {b[~ok][:3]}"


@pytest.mark.parametrize("test", ["wald", "lrt"])
def test_reference_golden_small(test):   # TS:719-756, TS:940-985
    hb = _hb()
    doc, x, y, cov = load_regression_logistic()
    mt = _mt(x, y=y, c1=cov[:, 1], c2=cov[:, 2])
    ht = hb.logistic_regression_rows(test=test, y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.c1, mt.c2])
    exp = doc["expected_" + test]
    for pos in ("1", "2"):
        for f, v in exp[pos].items():
            assert abs(ht[f][int(pos) - 1] - v) < 5e-7, (pos, f, ht[f][int(pos) - 1], v)
    for pos in exp["not_converged"]:
        assert not ht.fit["converged"][pos - 1]                       # separable
    for pos in exp["constant"]:
        i = pos - 1
        assert (not ht.fit["converged"][i]) or np.isnan(ht.p_value[i]) or abs(ht.p_value[i] - 1) < 1e-4
    r = ht.collect()[0]
    assert r.fit.converged and r.fit.n_iterations > 0 and not r.fit.exploded
    assert ht.row[-1] == "fit" and ht.row[-2] == "p_value"


def test_epacts_goldens_and_oracle():   # TS:1722-1862
    hb = _hb()
    z, x, y, cov = load_epacts()
    mt = _mt(x, is_case=y, is_female=cov[:, 1], pc1=cov[:, 2], pc2=cov[:, 3])
    covs = [1.0, mt.is_female, mt.pc1, mt.pc2]
    w = hb.logistic_regression_rows("wald", mt.is_case, mt.GT.n_alt_alleles(), covs)
    for i in range(5):
        for j, f in enumerate(("beta", "standard_error", "z_stat", "p_value")):
            assert w[f][i] == pytest.approx(z["wald"][i, j], rel=z["wald_rel"][i, j]), (i, f)
    lrt = hb.logistic_regression_rows("lrt", mt.is_case, mt.GT.n_alt_alleles(), covs)
    assert lrt.p_value == pytest.approx(z["lrt_p"], rel=1e-4)
    fi = hb.logistic_regression_rows("firth", mt.is_case, mt.GT.n_alt_alleles(), covs)
    assert fi.beta == pytest.approx(z["firth"][:, 0], rel=1e-4)
    assert fi.p_value == pytest.approx(z["firth"][:, 1], rel=1e-4)
    assert fi.fit["converged"].all() and w.fit["converged"].all()
    for test, ht in (("wald", w), ("lrt", lrt), ("firth", fi)):
        want = L.logreg_rows(test, x, y, cov)
        for f in ("beta", "standard_error", "z_stat", "chi_sq_stat", "p_value"):
            if f in want:
                _close(ht[f], want[f], 1e-6 if test != "firth" else 1e-5, f"{test} {f}")
        assert np.array_equal(ht.fit["n_iterations"], want["n_iterations"]), test
        assert np.array_equal(ht.fit["converged"], want["converged"]) and np.array_equal(ht.fit["exploded"], want["exploded"])


@pytest.mark.parametrize("test", ["wald", "lrt", "firth"])
@pytest.mark.parametrize("K", [1, 3, 6, 11, 14, 19, 24, 40, 63])
def test_vs_oracle_random(test, K):
    """Seeded binary phenotypes with a causal variant, missing calls / phenotypes / covariates, K covariates."""
    hb = _hb()
    N, M = 1200, 48 if test != "firth" else 24
    bn = hb.balding_nichols_model(3, N, M, missing_rate=0.04, seed=100 + K)
    x = bn.genotypes.to_dosage().astype(np.float64)
    x[x < 0] = np.nan
    x[1] = 0.0                                     # monomorphic: x is in the span of the intercept
    rng = np.random.default_rng(K)
    cov = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(K - 1)])
    eta = -0.3 + 0.7 * np.nan_to_num(x[5]) + (0.5 * cov[:, 1] if K > 1 else 0.0)
    y = (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(np.float64)
    y[rng.random(N) < 0.05] = np.nan
    if K > 1:
        cov[rng.random(N) < 0.02, K - 1] = np.nan
    mt = bn.annotate_cols(y=y, **{f"c{k}": cov[:, k] for k in range(K)})
    ht = hb.logistic_regression_rows(test, mt.y, mt.GT.n_alt_alleles(), [mt[f"c{k}"] for k in range(K)])
    want = L.logreg_rows(test, x, y, cov)
    same_fit = (ht.fit["converged"] == want["converged"]) & (ht.fit["n_iterations"] == want["n_iterations"])
    assert same_fit.mean() > 0.95, (test, K, int((~same_fit).sum()))    # a borderline |delta| ~ tol may take one step more
    conv = want["converged"] & ht.fit["converged"]
    assert conv.sum() >= M - 4
    rel = 2e-6 if test != "firth" else 2e-5
    for f in ("beta", "standard_error", "z_stat", "chi_sq_stat", "p_value"):
        if f in want:
            a, b = np.asarray(ht[f])[conv], want[f][conv]
            big = np.abs(b) > 1e-6 if f in ("beta", "z_stat", "chi_sq_stat") else np.ones_like(b, dtype=bool)
            _close(a[big], b[big], rel, f"{test} K={K} {f}")
    assert int(np.nanargmin(ht.p_value)) == 5                              # the causal variant


def test_multi_pheno_shapes_and_errors():
    hb = _hb()
    z = np.load(os.path.join(GOLDEN, "fastlmm.npz"))
    N, M = int(z["n_samples"]), 64
    rows = obed.bed_body(z["bed"], N, int(z["n_variants"]))[:M]
    x = obed.decode_rows(rows, N)
    rng = np.random.default_rng(3)
    y1 = (rng.random(N) < 0.4).astype(np.float64)
    y2 = (rng.random(N) < 0.6).astype(np.float64)
    c1 = rng.normal(size=N)
    mt = hb.MatrixTable(hb.PackedGenotypes.from_bed_rows(rows, N), rows={"rsid": np.arange(M)}, cols={"y1": y1, "y2": y2, "c1": c1})
    ht = hb.logistic_regression_rows("wald", [mt.y1, mt.y2], mt.GT.n_alt_alleles(), [1.0, mt.c1], pass_through=["rsid"])
    assert ht.beta.shape == (M, 2) and ht.fit["n_iterations"].shape == (M, 2) and list(ht.rsid) == list(range(M))
    r7 = ht.collect()[7]      # the reference's nested schema (TS:758-803): one struct per phenotype
    assert len(r7.logistic_regression) == 2
    for p in range(2):
        lr = r7.logistic_regression[p]
        assert (lr.beta == ht.beta[7, p] or np.isnan(lr.beta)) and lr.fit.n_iterations == ht.fit["n_iterations"][7, p]
        assert lr.fit.converged == bool(ht.fit["converged"][7, p])
    cov = np.column_stack([np.ones(N), c1])
    for col, yy in enumerate((y1, y2)):
        want = L.logreg_rows("wald", x, yy, cov)
        ok = want["converged"]
        _close(ht.beta[ok, col], want["beta"][ok], 2e-6, f"beta col {col}")
        _close(ht.p_value[ok, col], want["p_value"][ok], 2e-6, f"p col {col}")
    with pytest.raises(hb.FatalError, match="Failed to fit logistic regression null model"):   # TS:459-476
        hb.logistic_regression_rows("wald", mt.y1, mt.GT.n_alt_alleles(), [1.0, mt.c1], max_iterations=0)
    with pytest.raises(Exception, match="at most 63 covariates"):
        many = mt.annotate_cols(**{f"k{i}": rng.normal(size=N) for i in range(64)})
        hb.logistic_regression_rows("wald", many.y1, many.GT.n_alt_alleles(), [1.0] + [many[f"k{i}"] for i in range(64)])


@pytest.mark.parametrize("which", ["pl", "gp"])
def test_wald_on_dense_dosages(which):   # TS:851-938, through DenseDosage -> lrr_run_logit_dense
    hb = _hb()
    from tests.helpers import gp_dosage, pl_dosage
    doc, x, y, cov = load_regression_logistic()
    dos = pl_dosage(doc) if which == "pl" else gp_dosage(doc)
    mt = hb.MatrixTable(hb.DenseDosage(dos), cols={"y": y, "c1": cov[:, 1], "c2": cov[:, 2]})
    exp = doc["expected_wald_dosage"]
    tol = 5e-7 if which == "pl" else 5e-5
    ht = hb.logistic_regression_rows("wald", mt.y, mt.x, [1.0, mt.c1, mt.c2])
    for pos in ("1", "2"):
        for f, v in exp[pos].items():
            assert abs(ht[f][int(pos) - 1] - v) < tol, (which, pos, f, ht[f][int(pos) - 1])
    assert not ht.fit["converged"][2]
    for test in ("wald", "lrt", "firth"):
        got = hb.logistic_regression_rows(test, mt.y, mt.x, [1.0, mt.c1, mt.c2])
        want = L.logreg_rows(test, dos, y, cov)
        conv = want["converged"] & got.fit["converged"]
        info = (test, got.fit["converged"].tolist(), want["converged"].tolist(), got.fit["n_iterations"].tolist(),
                want["n_iterations"].tolist())
        assert np.array_equal(got.fit["converged"][:2], want["converged"][:2]), info
        assert conv[:2].all() or test == "firth", info
        for f in ("beta", "p_value"):
            _close(np.asarray(got[f])[:2][conv[:2]], want[f][:2][conv[:2]], 1e-5, f"{test} {f}")   # rows 6-10 are degenerate
    # the score test on the same dense dosages (lrr_run_score_dense; LogisticRegressionModel.scala:211-264)
    hs = hb.logistic_regression_rows("score", mt.y, mt.x, [1.0, mt.c1, mt.c2])
    ws = L.logreg_score(dos, y, cov)
    ok = np.isfinite(ws["chi_sq_stat"][:5])
    _close(hs.chi_sq_stat[:5][ok], ws["chi_sq_stat"][:5][ok], 1e-6, "score chi2 dense")
    _close(hs.p_value[:5][ok], ws["p_value"][:5][ok], 1e-5, "score p dense")


def test_score_on_dense_dosages_vs_packed_and_oracle():
    """The score test gives the same rows for a dense float64 field holding the hard calls as for the packed store, and
    matches the oracle on real-valued dosages with missing entries."""
    hb = _hb()
    rng = np.random.default_rng(8)
    N, M = 2500, 300
    mt = hb.balding_nichols_model(3, N, M, missing_rate=0.03, seed=19)
    dos = mt.genotypes.to_dosage().astype(np.float64)
    dos[dos < 0] = np.nan
    c1 = rng.normal(size=N)
    yb = (rng.random(N) < 1 / (1 + np.exp(-(0.4 * np.nan_to_num(dos[3]) - 0.3 + 0.2 * c1)))).astype(np.float64)
    mt = mt.annotate_cols(y=yb, c1=c1)
    packed = hb.logistic_regression_rows("score", mt.y, mt.GT.n_alt_alleles(), [1.0, mt.c1])
    dmt = hb.MatrixTable(hb.DenseDosage(dos), cols={"y": yb, "c1": c1})
    dense = hb.logistic_regression_rows("score", dmt.y, dmt.x, [1.0, dmt.c1])
    _close(dense.chi_sq_stat, packed.chi_sq_stat, 1e-9, "dense vs packed chi2")
    soft = np.clip(dos + 0.05 * rng.normal(size=dos.shape), 0, 2)
    smt = hb.MatrixTable(hb.DenseDosage(soft), cols={"y": yb, "c1": c1})
    got = hb.logistic_regression_rows("score", smt.y, smt.x, [1.0, smt.c1])
    want = L.logreg_score(soft, yb, np.column_stack([np.ones(N), c1]))
    _close(got.chi_sq_stat, want["chi_sq_stat"], 1e-6, "soft dosage chi2")
    _close(got.p_value, want["p_value"], 1e-5, "soft dosage p")


def test_logistic_at_400k_samples_vs_oracle():
    """The 256-thread float64 reductions of logit_fit_kernel and the score sweep at the BASELINE sample count."""
    hb = _hb()
    from hail_b200 import bn
    N, M, K = 400_000, 24, 4
    pop, th, _ = bn.bn_parameters(3, N, M, missing_rate=0.01, seed=23)
    gt = bn.bn_fill(hb.PackedGenotypes.empty(M, N), pop, th, seed=23)
    dos = gt.to_dosage().astype(np.float64)
    dos[dos < 0] = np.nan
    rng = np.random.default_rng(24)
    cov = np.column_stack([np.ones(N)] + [rng.standard_normal(N) for _ in range(K - 1)])
    eta = -0.5 + 0.03 * np.nan_to_num(dos[2]) + 0.2 * cov[:, 1]
    yb = (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(np.float64)
    yb[rng.random(N) < 0.02] = np.nan
    mt = hb.MatrixTable(gt, cols={"y": yb, **{f"c{i}": cov[:, i] for i in range(1, K)}})
    covs = [1.0] + [mt[f"c{i}"] for i in range(1, K)]
    for test in ("wald", "lrt", "score", "firth"):
        ht = hb.logistic_regression_rows(test, mt.y, mt.GT.n_alt_alleles(), covs)
        want = L.logreg_score(dos, yb, cov) if test == "score" else L.logreg_rows(test, dos, yb, cov)
        conv = np.ones(M, dtype=bool) if test == "score" else (want["converged"] & ht.fit["converged"])
        assert conv.sum() >= M - 1, test
        for f in ("beta", "standard_error", "z_stat", "chi_sq_stat", "p_value"):
            if f in want:
                a, b = np.asarray(ht[f])[conv], np.asarray(want[f])[conv]
                # (the LRT statistic is 2 (l1 - l0) with |l| ~ 2.7e5: below 1e-3 it is float64 roundoff in any implementation)
                floor = 1e-3 if f == "chi_sq_stat" else 1e-6
                big = np.abs(b) > floor if f in ("beta", "z_stat", "chi_sq_stat") else np.ones_like(b, dtype=bool)
                if f == "p_value" and test in ("lrt", "firth"):
                    big = np.asarray(want["chi_sq_stat"])[conv] > 1e-3
                _close(a[big], b[big], 2e-6 if test != "firth" else 2e-5, f"400k {test} {f}")
        assert int(np.nanargmin(ht.p_value)) == 2
