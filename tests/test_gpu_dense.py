"""GPU parity of the dense float64 `x` path (lrr_run_dense): `x` may be any entry-indexed float64 expression in the
reference (methods/statgen.py:229, 391), not only GT.n_alt_alleles() -- PL / GP dosages, imputed DS fields.

Mirrors test_statgen.py:286-316 (pl_dosage), :318-348 (gp_dosage), :350-364 (DS vs GT equivalence); the rest compares
the CUDA path with the CPU oracle on seeded inputs.
"""
import os

import numpy as np
import pytest

from oracle import linreg_oracle as O
from tests.helpers import assert_fields_close, gp_dosage, load_regression_linear, pl_dosage

pytestmark = pytest.mark.gpu


def _hb():
    import hail_b200 as hb
    return hb


def _dense_mt(x, **cols):
    hb = _hb()
    rows = {"locus": np.array([("1", i + 1) for i in range(x.shape[0])], dtype=object),
            "alleles": np.array([("C", "T")] * x.shape[0], dtype=object)}
    return hb.MatrixTable(hb.DenseDosage(x), rows=rows, cols=cols, row_key=("locus", "alleles"))


def _as_dict(ht):
    d = {"n": ht.n}
    for f in ("sum_x", "y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
        v = np.asarray(ht[f])
        d[f] = v if (v.ndim == 2 or f == "sum_x") else v[:, None]
    return d


@pytest.mark.parametrize("which", ["pl_dosage", "gp_dosage"])
def test_reference_golden_dosage(which):  # TS:286-316, TS:318-348
    hb = _hb()
    x, y, cov, doc = load_regression_linear()
    dos = pl_dosage(doc) if which == "pl_dosage" else gp_dosage(doc)
    mt = _dense_mt(dos, pheno=y, Cov1=cov[:, 0], Cov2=cov[:, 1])
    ht = hb.linear_regression_rows(y=mt.pheno, x=mt.x, covariates=[1.0, mt.Cov1, mt.Cov2])
    exp = doc["expected"][which]
    assert (ht.n == 6).all()
    for pos in ("1", "2", "3"):
        for f, v in exp[pos].items():
            tol = 5e-5 if (which == "gp_dosage" and f in ("beta", "standard_error")) else 5e-7
            assert abs(ht[f][int(pos) - 1] - v) < tol, (pos, f, ht[f][int(pos) - 1], v)
    assert np.isnan(ht.standard_error[5])     # TS:348
    want = O.linreg_group(dos, y[:, None], np.column_stack([np.ones(8), cov]))
    got = _as_dict(ht)
    good = slice(0, 5)   # rows 6-10 are degenerate (x in the covariate span): roundoff noise in the reference too
    assert_fields_close({k: v[good] for k, v in got.items()}, {k: v[good] for k, v in want.items() if k != "_d"},
                        t_floor=1e-9)


def test_ds_equals_gt():  # TS:350-364: a dosage field holding the hard calls gives the hard-call results
    hb = _hb()
    N, M = 1500, 200
    mt = hb.balding_nichols_model(3, N, M, missing_rate=0.03, seed=21)
    rng = np.random.default_rng(4)
    y = rng.normal(size=N)
    dos = mt.genotypes.to_dosage().astype(np.float64)
    dos[dos < 0] = np.nan
    mt = mt.annotate_cols(y=y)
    gt = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0], _kernel="fp64")
    dmt = _dense_mt(dos, y=y)
    ds = hb.linear_regression_rows(y=dmt.y, x=dmt.x, covariates=[1.0])
    assert np.array_equal(ds.n, gt.n)
    assert_fields_close(_as_dict(ds), _as_dict(gt), t_floor=1e-9)


@pytest.mark.parametrize("N,M,K,P,intercept", [(1000, 70, 3, 1, True), (2049, 33, 10, 3, True), (777, 130, 2, 2, False),
                                              (515, 40, 0, 1, False), (1200, 37, 4, 14, True)])
def test_dense_vs_oracle(N, M, K, P, intercept):
    """Float dosages in [0, 2] with NaN holes, missing phenotypes / covariates; C = K + P above and below one pass (12)."""
    hb = _hb()
    rng = np.random.default_rng(N + M)
    x = rng.beta(0.6, 1.4, size=(M, N)) * 2.0
    x[rng.random((M, N)) < 0.07] = np.nan
    cov = rng.normal(size=(N, K))
    if intercept and K:
        cov[:, 0] = 1.0
    ys = rng.normal(size=(N, P)) + 0.4 * np.nan_to_num(x[3])[:, None]
    ys[rng.random(N) < 0.05, 0] = np.nan
    if K:
        cov[rng.random(N) < 0.02, K - 1] = np.nan
    mt = _dense_mt(x, **{f"y{p}": ys[:, p] for p in range(P)}, **{f"c{k}": cov[:, k] for k in range(K)})
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ht = hb.linear_regression_rows(y=[mt[f"y{p}"] for p in range(P)], x=mt.x,
                                       covariates=[mt[f"c{k}"] for k in range(K)])
    want = O.linreg_group(x, ys, cov)
    assert_fields_close(_as_dict(ht), {k: v for k, v in want.items() if k != "_d"}, t_floor=1e-9)
    idx = O.complete_samples(ys, cov)[2]
    assert np.array_equal(ht.n_missing, np.isnan(x[:, idx]).sum(axis=1))


def test_dense_chained_and_edges():
    hb = _hb()
    rng = np.random.default_rng(9)
    N, M = 900, 45
    x = rng.random((M, N)) * 2.0
    x[rng.random((M, N)) < 0.25] = np.nan
    x[0] = np.nan            # all-missing variant: every statistic NaN (RU:52: 0 / 0)
    x[1] = 0.75              # constant column: degenerate
    y1 = rng.normal(size=N); y2 = rng.normal(size=N)
    y1[rng.random(N) < 0.1] = np.nan
    y2[rng.random(N) < 0.3] = np.nan
    c = rng.normal(size=N)
    mt = _dense_mt(x, y1=y1, y2=y2, c=c)
    ht = hb.linear_regression_rows(y=[[mt.y1], [mt.y2]], x=mt.x, covariates=[1.0, mt.c])
    cov = np.column_stack([np.ones(N), c])
    want = O.linreg_chained(x, [y1[:, None], y2[:, None]], cov)
    for g in range(2):
        got = {"n": ht.n[:, g], "sum_x": ht.sum_x[:, g]}
        for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
            got[f] = ht[f][g]
        assert np.isnan(got["sum_x"][0]) and np.isnan(got["beta"][0]).all() and np.isnan(got["p_value"][0]).all()
        assert np.isnan(got["standard_error"][1]).all() or (np.abs(got["t_stat"][1]) < 1e-3).all()
        assert_fields_close({k: v[2:] for k, v in got.items()}, {k: v[2:] for k, v in want[g].items() if k != "_d"},
                            t_floor=1e-9, ctx=f"g={g}")
    # empty row range; filtered columns
    e = _dense_mt(np.empty((0, N)), y=y1)
    assert hb.linear_regression_rows(y=e.y, x=e.x, covariates=[1.0]).count() == 0
    keep = rng.random(N) < 0.6
    sub = mt.filter_cols(keep)
    hs = hb.linear_regression_rows(y=sub.y1, x=sub.x, covariates=[1.0, sub.c])
    ws = O.linreg_group(x[:, keep], y1[keep][:, None], cov[keep])
    got = _as_dict(hs)
    assert_fields_close({k: v[2:] for k, v in got.items()}, {k: v[2:] for k, v in ws.items() if k != "_d"}, t_floor=1e-9)


def test_dense_rejects_mismatched_sources():
    hb = _hb()
    rng = np.random.default_rng(0)
    mt = _dense_mt(rng.random((4, 30)), y=rng.normal(size=30))
    with pytest.raises(hb.ExpressionException):
        mt.GT


@pytest.mark.parametrize("N,P,K", [(900, 2, 3), (1301, 1, 10)])
def test_dense_weighted_vs_oracle(N, P, K):
    """`weights` on a dense float64 x (statgen.py:557-581, 636-660): x is mean-imputed, then scaled by sqrt(w)."""
    hb = _hb()
    rng = np.random.default_rng(N)
    M = 60
    x = rng.beta(0.7, 1.3, size=(M, N)) * 2.0
    x[rng.random((M, N)) < 0.06] = np.nan
    cov = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(K - 1)])
    ys = rng.normal(size=(N, P)) + 0.3 * np.nan_to_num(x[2])[:, None]
    ys[rng.random(N) < 0.04, 0] = np.nan
    w = rng.uniform(0.2, 4.0, size=N)
    w[rng.random(N) < 0.03] = np.nan                    # missing weights drop the sample (TS:610-660)
    mt = _dense_mt(x, w=w, **{f"y{p}": ys[:, p] for p in range(P)}, **{f"c{k}": cov[:, k] for k in range(1, K)})
    covs = [1.0] + [mt[f"c{k}"] for k in range(1, K)]
    ht = hb.linear_regression_rows(y=[mt[f"y{p}"] for p in range(P)], x=mt.x, covariates=covs, weights=mt.w)
    want = O.linreg_group_weighted(x, ys, cov, w)
    assert_fields_close(_as_dict(ht), {k: v for k, v in want.items() if k != "_d"}, t_floor=1e-9, ctx="dense weighted")
    # unit weights reproduce the unweighted statistics
    mt1 = mt.annotate_cols(one=np.ones(N))
    hu = hb.linear_regression_rows(y=mt1.y0, x=mt1.x, covariates=[1.0] + [mt1[f"c{k}"] for k in range(1, K)], weights=mt1.one)
    h0 = hb.linear_regression_rows(y=mt1.y0, x=mt1.x, covariates=[1.0] + [mt1[f"c{k}"] for k in range(1, K)])
    assert np.allclose(hu.beta, h0.beta, rtol=1e-7) and np.allclose(hu.p_value, h0.p_value, rtol=1e-6)


def test_lambda_gc():   # statgen.py:3096-3128
    hb = _hb()
    import scipy.stats as st
    rng = np.random.default_rng(8)
    p = rng.random(10001)
    p[::50] = np.nan
    p[3] = 1e-300
    ok = p[~np.isnan(p)]
    want = np.median(st.chi2.isf(ok, 1)) / st.chi2.isf(0.5, 1)
    assert hb.lambda_gc(p) == pytest.approx(want, rel=1e-10)
    assert hb.lambda_gc(p[:-1]) == pytest.approx(np.median(st.chi2.isf(p[:-1][~np.isnan(p[:-1])], 1)) / st.chi2.isf(0.5, 1), rel=1e-10)
    assert np.isnan(hb.lambda_gc(np.array([np.nan])))
    # uniform p-values: no inflation
    assert abs(hb.lambda_gc(rng.random(200001)) - 1.0) < 0.02


def test_dense_infinite_entries_follow_the_reference():
    """An infinite entry is a DEFINED value in the reference (RU:16-58 only imputes missing ones): sum_x = +-Inf (NaN when
    both signs occur), x.x = Inf and every statistic NaN (Inf - Inf).  The other rows of the same CTA are untouched, and
    an infinite entry of a sample OUTSIDE the group (missing phenotype) changes nothing."""
    hb = _hb()
    rng = np.random.default_rng(31)
    N, M = 900, 70
    x = rng.uniform(0, 2, size=(M, N))
    x[rng.random(x.shape) < 0.02] = np.nan
    y = rng.normal(size=N)
    y[5] = np.nan                      # sample 5 is outside the group
    bmt = _dense_mt(x, y=y, c=np.arange(N) / N)
    base = hb.linear_regression_rows(y=bmt.y, x=bmt.x, covariates=[1.0, bmt.c])
    xi = x.copy()
    xi[3, 17] = np.inf
    xi[9, 100] = -np.inf
    xi[20, 40], xi[20, 41] = np.inf, -np.inf
    xi[33, 5] = np.inf                 # outside the group: ignored
    mt = _dense_mt(xi, y=y, c=np.arange(N) / N)
    ht = hb.linear_regression_rows(y=mt.y, x=mt.x, covariates=[1.0, mt.c])
    assert ht.sum_x[3] == np.inf and ht.sum_x[9] == -np.inf and np.isnan(ht.sum_x[20])
    for row in (3, 9, 20):
        for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
            assert np.isnan(ht[f][row]), (row, f)
    assert np.array_equal(ht.n_missing, np.isnan(x[:, np.arange(N) != 5]).sum(axis=1))   # Inf is not a missing entry
    others = np.setdiff1d(np.arange(M), [3, 9, 20])
    for f in ("sum_x", "y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
        assert np.array_equal(ht[f][others], base[f][others], equal_nan=True), f
    # what the reference's float64 algebra gives for such a row (oracle, same semantics)
    want = O.linreg_group(xi, y[:, None], np.column_stack([np.ones(N), np.arange(N) / N]))
    assert np.isnan(want["beta"][[3, 9, 20], 0]).all() and want["sum_x"][3] == np.inf


@pytest.mark.parametrize("kind", ["bgen8", "float"])
def test_compact_u16_dosages_equal_the_dense_path_bit_for_bit(kind):
    """the 8 / 16-bit dosage store (lrr_run_dense_u16).  The regression of a CompactDosage must equal the
    float64 dense path on the DEQUANTISED values bit for bit (same kernel, the entries are converted in shared memory), and
    the oracle within tolerance: BGEN-style dosages (multiples of 1/255 from 8-bit probabilities), general floats, missing
    entries, a sample count that is no multiple of 8, a sample outside the group, 14 phenotypes (two column passes)."""
    hb = _hb()
    rng = np.random.default_rng(41)
    N, M, P = 1003, 150, 14
    if kind == "bgen8":
        p1, p2 = rng.integers(0, 256, size=(M, N)), rng.integers(0, 256, size=(M, N))
        p2 = np.minimum(p2, 255 - p1)
        x = (p1 + 2.0 * p2) / 255.0
    else:
        x = rng.uniform(0, 2, size=(M, N))
    x[rng.random(x.shape) < 0.03] = np.nan
    x[7] = np.nan                                           # an all-missing row
    ys = rng.normal(size=(N, P)) + 0.3 * np.nan_to_num(x[:P].T)
    ys[11, :] = np.nan                                      # sample 11 is outside the group
    c = rng.normal(size=N)
    cd = hb.CompactDosage(x)
    assert cd.scale == (1.0 / 255.0 if kind == "bgen8" else 2.0 / 65534.0) and cd.nbytes == M * 1008 * 2
    deq = cd.to_dosage()
    assert np.array_equal(np.isnan(deq), np.isnan(x)) and np.nanmax(np.abs(deq - x)) <= (1e-12 if kind == "bgen8" else 1.6e-5)
    cols = {**{f"y{i}": ys[:, i] for i in range(P)}, "c": c}
    cmt = hb.MatrixTable(cd, cols=cols)
    dmt = hb.MatrixTable(hb.DenseDosage(deq), cols=cols)
    hc = hb.linear_regression_rows(y=[cmt[f"y{i}"] for i in range(P)], x=cmt.x, covariates=[1.0, cmt.c])
    hd = hb.linear_regression_rows(y=[dmt[f"y{i}"] for i in range(P)], x=dmt.x, covariates=[1.0, dmt.c])
    for f in ("n", "sum_x", "y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
        assert np.array_equal(hc[f], hd[f], equal_nan=True), f
    assert np.array_equal(hc.n_missing, hd.n_missing) and hc.n_missing[7] == N - 1
    want = O.linreg_group(deq, ys, np.column_stack([np.ones(N), c]))
    good = np.arange(M) != 7
    assert_fields_close({k: v[good] for k, v in _as_dict(hc).items()}, {k: v[good] for k, v in want.items() if k != "_d"}, t_floor=1e-9)
    # chained groups on the same store
    h2 = hb.linear_regression_rows(y=[[cmt.y0], [cmt.y1, cmt.y2]], x=cmt.x, covariates=[1.0, cmt.c])
    # (the host prologue of a group of 1 or 2 phenotypes rounds differently from that of 14: last-bit agreement only)
    assert np.allclose(h2.beta[0][:, 0], hc.beta[:, 0], rtol=1e-10, equal_nan=True) and np.allclose(h2.beta[1], hc.beta[:, 1:3], rtol=1e-10, equal_nan=True)


def test_import_bgen_to_regression(tmp_path):
    """SURVEY 8f rank 3, the file side: `import_bgen` on the reference's own BGEN resource (first 64 variants of
    example.8bits.bgen, tests/golden/bgen_example.npz) gives the compact dosage store; the regression on `mt.dosage` equals
    the oracle on the dequantised dosages, and the dosages are the text file's (example.gen) to the format's 8 bits
    (test_impex.py:1264-1272)."""
    hb = _hb()
    from hail_b200.impex import import_bgen
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bgen_example.npz"))
    path = tmp_path / "example64.bgen"
    path.write_bytes(z["bgen"].tobytes())
    with pytest.raises(hb.FatalError, match="Invalid locus '01:2000'"):
        import_bgen(str(path), entry_fields=["dosage"])                      # contig '01' is not a GRCh37 contig
    mt = import_bgen(str(path), entry_fields=["dosage"], contig_recoding={"01": "1"})
    assert mt.count() == (64, 500) and list(mt.s) == list(z["samples"])
    assert tuple(mt.row["locus"][0]) == ("1", 2000) and tuple(mt.row["alleles"][0]) == ("A", "G") and mt.row["rsid"][0] == "RSID_2"
    deq = mt.genotypes.to_dosage()
    assert np.array_equal(np.isnan(deq), np.isnan(z["gen_dosage"])) and np.nanmax(np.abs(deq - z["gen_dosage"])) <= 3.0 / 255 + 1e-6
    rng = np.random.default_rng(8)
    N = 500
    c = rng.normal(size=N)
    y = rng.normal(size=N) + 0.4 * np.nan_to_num(deq[3])
    mt = mt.annotate_cols(y=y, c=c)
    ht = hb.linear_regression_rows(y=mt.y, x=mt.dosage, covariates=[1.0, mt.c])
    want = O.linreg_group(deq, y[:, None], np.column_stack([np.ones(N), c]))
    assert_fields_close(_as_dict(ht), {k: v for k, v in want.items() if k != "_d"}, t_floor=1e-9, ctx="bgen")
    assert int(np.nanargmin(ht.p_value)) == 3
