"""The tolerance is enforced, not hoped for: BASELINE.json's own shapes and the inputs that stress the quantised sweeps.

* C3 exactly: 400k samples, 25 % missing calls, chained y=[[y1],[y2]] with 10 % / 20 % phenotype missingness, K = 10
  (LinearRegression.scala:296-347: per-group imputation and algebra) against the C oracle.
* C4 exactly: 400k samples x 128 phenotypes x 10 covariates against the C oracle.
* Heavy-tailed and un-centred covariates (outliers 1e3x / 1e6x the typical entry, no-intercept models with mean >> sd):
  whatever the digit quantisation costs must be caught by the per-variant bound (include/lrr_b200.h lrr_set_guard) and
  repaired by the float64 recompute -- AUTO never returns a row outside rel 1e-6 / 1e-5.
* The f32 accumulators of the 4-bit sweep driven to their exactness bound 2^24 (constant-sign digit 8 on every sample
  x all-missing / all-hom-alt rows) at 400k, 500k and 699k samples, and the refusal above that bound.

Tolerances: BASELINE.json's (tests/helpers.assert_fields_close).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import bed as obed
from oracle import linreg_oracle as O
from tests.helpers import assert_fields_close, ytx_floor_of
from tests.test_gpu_parity import _as_oracle_dict, _big_case, _hb


def _strip(want):
    return {k: v for k, v in want.items() if k != "_d"}


def _ctx():
    from hail_b200 import _lib
    return _lib.context(0)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", ["auto", "tc4", "tc"])
def test_c3_at_its_own_shape(kernel):
    """BASELINE config 3: 400k samples, 25 % missing, chained [[y1],[y2]] with per-group missingness, 782 chunks per row."""
    hb = _hb()
    from oracle import c_oracle
    N, M, K = 400_000, 1024, 10
    gt, bed_rows, cov, rng = _big_case(N, M, K, 0.25, seed=31)
    dos0 = obed.decode_rows(bed_rows[:1], N)[0]
    y1 = rng.standard_normal(N) + 0.02 * np.nan_to_num(dos0)
    y2 = rng.standard_normal(N)
    y1[rng.random(N) < 0.10] = np.nan
    y2[rng.random(N) < 0.20] = np.nan
    mt = hb.MatrixTable(gt, cols={"y1": y1, "y2": y2, **{f"c{i}": cov[:, i] for i in range(1, K)}})
    ht = hb.linear_regression_rows(y=[[mt.y1], [mt.y2]], x=mt.GT.n_alt_alleles(),
                                   covariates=[1.0] + [mt[f"c{i}"] for i in range(1, K)], _kernel=kernel)
    if kernel == "auto":
        assert _ctx().last_kernel == "tc4"
    for g, y in enumerate((y1, y2)):
        want = c_oracle.linreg_group_bed(bed_rows, N, y[:, None], cov)
        got = {"n": ht.n[:, g], "sum_x": ht.sum_x[:, g]}
        for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
            got[f] = ht[f][g]
        assert_fields_close(got, _strip(want), t_floor=1e-9, ctx=f"C3 {kernel} group {g}")
        idx = O.complete_samples(y[:, None], cov)[2]
        x = obed.decode_rows(bed_rows[:64], N)
        assert np.array_equal(ht.n_missing[g][:64], np.isnan(x[:, idx]).sum(axis=1))


@pytest.mark.parametrize("kernel", ["auto", "tc"])
def test_c4_at_its_own_shape(kernel):
    """BASELINE config 4: 400k samples x 128 phenotypes x 10 covariates (the multi-pass dense contraction)."""
    hb = _hb()
    from oracle import c_oracle
    N, M, K, P = 400_000, 256, 10, 128
    gt, bed_rows, cov, rng = _big_case(N, M, K, 0.0, seed=37)
    dos = obed.decode_rows(bed_rows[:4], N)
    ys = rng.standard_normal((N, P))
    ys[:, 0] += 0.02 * dos[0]
    ys[:, 5] += 0.01 * dos[3]
    mt = hb.MatrixTable(gt, cols={**{f"y{i}": ys[:, i] for i in range(P)}, **{f"c{i}": cov[:, i] for i in range(1, K)}})
    ht = hb.linear_regression_rows(y=[mt[f"y{i}"] for i in range(P)], x=mt.GT.n_alt_alleles(),
                                   covariates=[1.0] + [mt[f"c{i}"] for i in range(1, K)], _kernel=kernel)
    if kernel == "auto":
        assert _ctx().last_kernel == "tc4"
        # the digit policy is part of the contract (it silently cost a seventh pass for most of round 2): 128 ten-digit
        # phenotype columns + 9 twelve-digit covariate columns + one "ones" row per pass fill SIX 240-column passes
        launches, mma_cols, _, digit_cols = _ctx().last_sweep_shape
        assert (launches, mma_cols, digit_cols) == (6, 1440, 128 * 10 + 9 * 12 + 6), _ctx().last_sweep_shape
    want = c_oracle.linreg_group_bed(bed_rows, N, ys, cov)
    assert ht.beta.shape == (M, P)
    # P > 2 is the many-phenotype profile: 10-digit phenotype columns, so a y_transpose_x that cancels to ~1e-5 of its scale
    # is held to that profile's stated floor (5e-7 standard errors of x.y) instead of 1e-6 of itself
    floor = None
    if kernel != "tc":
        floor = ytx_floor_of(obed.decode_rows(bed_rows, N).astype(np.float64), cov, want["standard_error"])
    assert_fields_close(_as_oracle_dict(ht), _strip(want), t_floor=1e-9, ctx=f"C4 {kernel}", ytx_abs=floor)
    assert int(np.nanargmin(ht.p_value[:, 0])) == 0


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("missing_rate", [0.0, 0.02])
def test_centred_bound_at_extreme_allele_frequencies(missing_rate):
    """Rows at allele frequency 0.5, ~1 and ~0 (sum x up to 2 n next to a small x.x - |Q'x|^2): the 4-bit sweep's 6-digit
    covariate columns stay inside the tolerance because the epilogue centres every dot product on the row's prevailing call
    (adds ac * sum of the basis rounding errors, bounds the rest by (quantum / 2) sum |x - ac|); such rows must neither
    leave the tolerance nor be sent to the float64 recompute wholesale."""
    hb = _hb()
    from oracle import c_oracle
    N, M, K = 400_000, 512, 10
    rng = np.random.Generator(np.random.Philox(key=[91, 3]))
    p = np.concatenate([np.full(64, 0.5), rng.uniform(0.95, 0.9995, 192), rng.uniform(0.0005, 0.05, 192),
                        rng.uniform(0.3, 0.7, 64)])
    x = np.empty((M, N), dtype=np.int8)
    for lo in range(0, M, 64):
        x[lo:lo + 64] = rng.binomial(2, p[lo:lo + 64, None], size=(min(64, M - lo), N)).astype(np.int8)
    if missing_rate > 0:
        for lo in range(0, M, 64):
            x[lo:lo + 64][rng.random((min(64, M - lo), N)) < missing_rate] = -1
    gt = hb.PackedGenotypes.from_dosage(x)
    bed_rows = obed.encode_rows(x)
    cov = np.column_stack([np.ones(N)] + [rng.standard_normal(N) for _ in range(K - 1)])
    y = rng.standard_normal(N) + 0.05 * (x[300] == 1)
    mt = hb.MatrixTable(gt, cols={"y": y, **{f"c{i}": cov[:, i] for i in range(1, K)}})
    covs = [1.0] + [mt[f"c{i}"] for i in range(1, K)]
    want = c_oracle.linreg_group_bed(bed_rows, N, y[:, None], cov)
    ht = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=covs, _kernel="tc4")
    assert_fields_close(_as_oracle_dict(ht), _strip(want), t_floor=1e-9, ctx=f"extreme AF miss={missing_rate}")
    # a handful of rows whose y_transpose_x cancels may be listed; the high-frequency rows as a class must not be
    assert _ctx().last_recomputed <= 8, _ctx().last_recomputed
    # and with the guard off the centred dot products alone are already inside the tolerance
    _ctx().check(_ctx().lib.lrr_set_guard(_ctx().handle, 0))
    try:
        raw = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=covs, _kernel="tc4")
    finally:
        _ctx().check(_ctx().lib.lrr_set_guard(_ctx().handle, 1))
    rawd = _as_oracle_dict(raw)
    for f in ("beta", "standard_error", "t_stat"):
        w = np.asarray(want[f]).reshape(-1)
        assert O.d_eq(np.asarray(rawd[f]).reshape(-1), w, 1e-6)[np.isfinite(w)].mean() > 0.99, f


# ---------------------------------------------------------------------------------------------
def _cov_case(name, N, rng):
    """Covariate matrices that stress a fixed-point basis; returns (cov [N, K], covariate list builder flag)."""
    z = rng.standard_normal((N, 5))
    if name == "outlier_1e3":
        z[17, 1] *= 1e3
        z[N // 2, 3] = -2e3
        return np.column_stack([np.ones(N), z]), True
    if name == "outlier_1e6":
        z[5, 0] = 1e6
        z[N - 3, 4] = -3e6
        return np.column_stack([np.ones(N), z]), True
    if name == "cauchy":
        return np.column_stack([np.ones(N), rng.standard_cauchy((N, 4))]), True
    if name == "uncentred_no_intercept":     # mean >> sd and no constant in the model
        return np.column_stack([1000.0 + z[:, 0], 50.0 + 0.01 * z[:, 1], z[:, 2] - 300.0]), False
    if name == "uncentred_with_intercept":
        return np.column_stack([np.ones(N), 1e4 + z[:, 0], 1e-3 * z[:, 1] + 7.0, z[:, 2]]), True
    raise KeyError(name)


@pytest.mark.parametrize("case", ["outlier_1e3", "outlier_1e6", "cauchy", "uncentred_no_intercept", "uncentred_with_intercept"])
@pytest.mark.parametrize("kernel", ["auto", "tc4", "tc"])
def test_heavy_tailed_and_uncentred_covariates(case, kernel):
    hb = _hb()
    from oracle import c_oracle
    N, M = 120_000, 8192          # large enough for AUTO to take the tensor-core sweep
    gt, bed_rows, _, rng = _big_case(N, M, 2, 0.01, seed=41)
    cov, has_const = _cov_case(case, N, rng)
    K = cov.shape[1]
    dos0 = obed.decode_rows(bed_rows[:1], N)[0]
    y = rng.standard_normal(N) + 0.05 * np.nan_to_num(dos0) + 0.3 * cov[:, -1] / max(1.0, np.abs(cov[:, -1]).max() ** 0.5)
    y[rng.random(N) < 0.02] = np.nan
    first = 1 if has_const else 0
    mt = hb.MatrixTable(gt, cols={"y": y, **{f"c{i}": cov[:, i] for i in range(first, K)}})
    covs = ([1.0] if has_const else []) + [mt[f"c{i}"] for i in range(first, K)]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ht = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=covs, _kernel=kernel)
    ctx = _ctx()
    if kernel == "auto":
        assert ctx.last_kernel == "tc4"
    recomputed = ctx.last_recomputed
    want = c_oracle.linreg_group_bed(bed_rows, N, y[:, None], cov)
    got = _as_oracle_dict(ht)
    nondeg = np.isfinite(want["standard_error"]).all(axis=1)
    assert nondeg.sum() > 0.95 * M
    assert_fields_close({k: v[nondeg] for k, v in got.items()}, {k: v[nondeg] for k, v in _strip(want).items()},
                        t_floor=1e-9, ctx=f"{case} {kernel} (recomputed {recomputed} of {M})")


def test_guard_recomputes_collinear_rows_and_reports_them():
    """Rows that are almost in the span of the covariates (x.x - |Q'x|^2 cancels to ~1e-9 of x.x) cannot be served by
    a fixed-point basis: the bound must list them, and the float64 recompute must agree with the oracle."""
    hb = _hb()
    from oracle import c_oracle
    N, M = 100_000, 20_480
    gt, bed_rows, _, rng = _big_case(N, M, 2, 0.0, seed=43)
    x = obed.decode_rows(bed_rows[:8], N)
    # covariates 1..4 reproduce variants 0..3 up to a small perturbation: those rows are nearly collinear
    cov = np.column_stack([np.ones(N)] + [x[i] + 1e-4 * rng.standard_normal(N) for i in range(4)] + [rng.standard_normal(N)])
    y = rng.standard_normal(N)
    mt = hb.MatrixTable(gt, cols={"y": y, **{f"c{i}": cov[:, i] for i in range(1, 6)}})
    ht = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0] + [mt[f"c{i}"] for i in range(1, 6)],
                                   _kernel="tc4")
    ctx = _ctx()
    n_re = ctx.last_recomputed
    assert n_re >= 4, n_re
    want = c_oracle.linreg_group_bed(bed_rows, N, y[:, None], cov)
    # the collinear rows amplify float64 roundoff by 1e8 in ANY implementation: compare them at 1e-3, the rest at 1e-6
    rest = np.arange(M) >= 4

    def check(table):
        got = _as_oracle_dict(table)
        assert_fields_close({k: v[rest] for k, v in got.items()}, {k: v[rest] for k, v in _strip(want).items()}, t_floor=1e-9)
        assert np.allclose(got["beta"][:4], want["beta"][:4], rtol=1e-3) and np.allclose(got["t_stat"][:4], want["t_stat"][:4], rtol=1e-3)

    check(ht)
    # genotype-like covariates are STRUCTURED (|Q'x| is large for every correlated row): the precision pilot of lrr_run
    # (first 8,192 rows) sees more than 2 % of its rows leave the bound and re-quantises the covariate columns with more
    # digits, so that far fewer than half of the rows need the float64 recompute
    assert n_re < M // 2, n_re
    # guard off: nothing is listed, the raw quantised rows come back (and the collinear ones are measurably off)
    raw = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0] + [mt[f"c{i}"] for i in range(1, 6)],
                                    _kernel="tc4", _guard=False)
    assert ctx.last_recomputed == 0
    assert np.isfinite(raw.t_stat[4:]).all()


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [400_000, 500_000, 699_000])
def test_f32_accumulators_exact_at_the_bound(N):
    """Every sample carries the digit of largest magnitude (8) with the same sign in one digit column, and the rows are
    all-missing / all-hom-alt / mostly-missing: |sum c u| reaches 24 (N - 1) (within 0.2 % of 2^24 at 699,000 samples).
    The sums must still be exact: results equal the float64 kernel's to roundoff."""
    hb = _hb()
    M = 256
    x = np.empty((M, N), dtype=np.int8)
    rng = np.random.default_rng(N)
    x[0::4] = -1                                            # all missing
    x[1::4] = 2                                             # all hom-alt
    x[2::4] = -1
    x[2::4, :50] = rng.integers(0, 3, size=(M // 4, 50))    # mostly missing, mean defined
    x[3::4] = rng.integers(0, 3, size=(M // 4, N))
    gt = hb.PackedGenotypes.from_dosage(x)
    # phenotype: one sample at 1.0 (the column maximum), every other at 8 quanta of the 13-digit fixed point: I = 8 -> lowest digit 8
    imax = np.floor(13.0 ** 13 / 3.0)
    y = np.full(N, 8.0 / imax)
    y[0] = 1.0
    mt = hb.MatrixTable(gt, cols={"y": y})
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # guard off: y sits at the quantum of its own fixed point, so the bound would (rightly) send every row to the
        # float64 recompute -- this test is about the raw f32 accumulators
        h4 = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[], _kernel="tc4", _guard=False)
        assert _ctx().last_recomputed == 0
        h64 = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[], _kernel="fp64")
    assert np.array_equal(h4.n_missing, (x < 0).sum(axis=1))
    assert np.array_equal(h4.n_missing, h64.n_missing)
    assert np.array_equal(h4.sum_x, h64.sum_x, equal_nan=True)
    ok = np.isfinite(h64.y_transpose_x)
    assert ok.sum() >= M // 2
    # y_transpose_x = sum_j x_j y_j with y an exact multiple of the quantum: the 4-bit sweep is EXACT, float64 FMA is not
    xf = np.where(x < 0, np.nan, x).astype(np.float64)
    mean = np.nanmean(xf[:, :], axis=1)
    exact = np.array([np.dot(np.where(np.isnan(r), m, r), y) if np.isfinite(m) else np.nan for r, m in zip(xf[:8], mean[:8])])
    assert np.allclose(h4.y_transpose_x[:8][np.isfinite(exact)], exact[np.isfinite(exact)], rtol=1e-12)
    assert np.allclose(h4.y_transpose_x[ok], h64.y_transpose_x[ok], rtol=1e-11)


def test_f32_accumulator_bound_refuses_above_2_pow_24():
    """699,100 samples with the same worst-case column: 24 * 699,099 > 2^24 -> the 4-bit sweep refuses, AUTO falls back
    to the int8 sweep (INT32 sums) and stays exact."""
    hb = _hb()
    from hail_b200._lib import LrrError
    N, M = 699_100, 3072      # large enough for AUTO to leave the float64 kernel
    rng = np.random.default_rng(7)
    x = np.tile(rng.integers(0, 3, size=(256, N), dtype=np.int8), (M // 256, 1))
    x[0] = -1
    x[0, :10] = 1
    gt = hb.PackedGenotypes.from_dosage(x)
    imax = np.floor(13.0 ** 13 / 3.0)
    y = np.full(N, 8.0 / imax)
    y[0] = 1.0
    mt = hb.MatrixTable(gt, cols={"y": y})
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(LrrError, match="exact range"):
            hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[], _kernel="tc4")
        ha = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[], _kernel="auto")
        assert _ctx().last_kernel == "tc"
        h64 = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[], _kernel="fp64")
    assert np.array_equal(ha.n_missing, h64.n_missing) and np.array_equal(ha.sum_x, h64.sum_x)
    assert np.allclose(ha.y_transpose_x, h64.y_transpose_x, rtol=1e-11)


def test_environment_cannot_change_results(monkeypatch):
    """The shipped library never reads the environment (the round-1 ablation switches are compiled out)."""
    hb = _hb()
    N, M = 3000, 700
    mt = hb.balding_nichols_model(3, N, M, missing_rate=0.02, seed=5)
    rng = np.random.default_rng(1)
    mt = mt.annotate_cols(y=rng.standard_normal(N), c=rng.standard_normal(N))
    base = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.c], _kernel="tc4")
    for var in ("LRR_ABL_CONTIG", "LRR_ABL_STREAM", "LRR_ABL_BITS", "LRR_TC_CLUSTER", "LRR_TC4_NU1", "LRR_ABL_L2PNONE"):
        monkeypatch.setenv(var, "1")
    again = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.c], _kernel="tc4")
    for f in ("beta", "standard_error", "t_stat", "p_value", "sum_x"):
        assert np.array_equal(base[f], again[f], equal_nan=True), f


def test_structured_covariates_streamed_from_host_blocks():
    """The precision pilot also runs when the rows arrive block by block (lrr_stream_*: every block is a short lrr_run):
    genotype-like covariates from host-resident .bed bytes, results within tolerance of the oracle."""
    hb = _hb()
    from oracle import c_oracle
    N, M = 100_000, 6144
    gt, bed_rows, _, rng = _big_case(N, M, 2, 0.0, seed=47)
    x = obed.decode_rows(bed_rows[:4], N)
    cov = np.column_stack([np.ones(N)] + [x[i] + 0.3 * rng.standard_normal(N) for i in range(3)] + [rng.standard_normal(N)])
    y = rng.standard_normal(N)
    host = hb.MatrixTable(hb.HostBedGenotypes(bed_rows, N), cols={"y": y, **{f"c{i}": cov[:, i] for i in range(1, 5)}})
    ht = hb.linear_regression_rows(y=host.y, x=host.GT.n_alt_alleles(), covariates=[1.0] + [host[f"c{i}"] for i in range(1, 5)],
                                   _stream_block=4096)   # (blocks large enough for AUTO to take the tensor-core sweep)
    want = c_oracle.linreg_group_bed(bed_rows, N, y[:, None], cov)
    assert_fields_close(_as_oracle_dict(ht), _strip(want), t_floor=1e-9, ctx="streamed, structured covariates")
