"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Mirrors hail/python/test/hail/methods/test_statgen.py:62-93 (basic), 134-221 (chained), 223-284 (R goldens),
366-457 (fam / multi-pheno).  Tolerances are BASELINE.json's: n / n_missing / integer sums exact; sum_x,
y_transpose_x, beta, standard_error, t_stat relative 1e-6 (the reference's own `_same` comparator); p_value
relative 1e-5.  `t_floor=1e-9` additionally accepts |delta t| <= 1e-9 for statistics whose true value is ~0
(relative error of a float64 dot product that cancels to ~0 is unbounded in any implementation).
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import bed as obed
from oracle import linreg_oracle as O
from tests.bn_mirror import bn_fill_numpy
from tests.helpers import GOLDEN, assert_fields_close, load_regression_linear, ytx_floor_of

KERNELS = ["fp64", "tc", "tc4"]
TC_KERNELS = ["tc", "tc4"]   # the exact-integer tensor-core sweeps: int8 digits / INT32 sums, E2M1 digits / f32 sums


def _hb():
    import hail_b200 as hb
    return hb


def _kernel_available(kernel, hb, n_groups=1):
    return True


def _mt_from_dosage(x, **cols):
    hb = _hb()
    gt = hb.PackedGenotypes.from_dosage(np.where(np.isnan(x), -1, x).astype(np.int8))
    rows = {"locus": np.array([("1", i + 1) for i in range(x.shape[0])], dtype=object),
            "alleles": np.array([("C", "T")] * x.shape[0], dtype=object),
            "qual": np.arange(x.shape[0], dtype=np.float64)}
    return hb.MatrixTable(gt, rows=rows, cols=cols, row_key=("locus", "alleles"))


def _as_oracle_dict(ht, P_list=False):
    d = {"n": ht.n}
    for f in ("sum_x", "y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
        v = np.asarray(ht[f])
        d[f] = v if (v.ndim == 2 or f == "sum_x") else v[:, None]
    return d


# ---------------------------------------------------------------------------------------------
def test_pack_roundtrips_and_bed_codes():
    hb = _hb()
    rng = np.random.default_rng(0)
    for N in (1, 5, 16, 17, 250, 513, 1031):
        x = rng.integers(-1, 3, size=(37, N)).astype(np.int8)
        g = hb.PackedGenotypes.from_dosage(x)
        assert g.stride % 128 == 0
        assert np.array_equal(g.to_dosage(), x)
        xf = np.where(x < 0, np.nan, x).astype(np.float64)
        rows = obed.encode_rows(xf)
        g2 = hb.PackedGenotypes.from_bed_rows(rows, N)
        assert np.array_equal(g2.to_dosage(), x)
        assert torch.equal(g.data, g2.data)  # both packers produce the identical store, padding included


def test_pack_reference_plink_fixtures():
    hb = _hb()
    for name in ("fastlmm.npz", "bn_4x1024.npz"):
        z = np.load(os.path.join(GOLDEN, name))
        N, M = int(z["n_samples"]), int(z["n_variants"])
        rows = obed.bed_body(z["bed"], N, M)
        want = obed.decode_rows(rows, N)
        got = hb.PackedGenotypes.from_bed_rows(rows, N).to_dosage().astype(np.float64)
        got[got < 0] = np.nan
        assert np.array_equal(got, want, equal_nan=True)


def test_bn_fill_matches_numpy_mirror():
    hb = _hb()
    from hail_b200 import bn
    for (N, M, mr) in ((100, 64, 0.0), (1000, 300, 0.25)):
        pop, th, af = bn.bn_parameters(3, N, M, missing_rate=mr, seed=7)
        g = bn.bn_fill(hb.PackedGenotypes.empty(M, N), pop, th, seed=7)
        want = bn_fill_numpy(th, pop, N, seed=7)
        got = g.to_dosage()
        assert np.array_equal(got, want)
        # shard regeneration: rows [100, 164) generated on their own equal the slice
        if M > 164:
            pop2, th2, _ = bn.bn_parameters(3, N, 64, missing_rate=mr, seed=7, first_variant=100)
            g2 = bn.bn_fill(hb.PackedGenotypes.empty(64, N), pop2, th2, seed=7, first_variant=100)
            assert np.array_equal(g2.to_dosage(), want[100:164])
        if mr:
            assert abs((want < 0).mean() - mr) < 0.01
        obs = np.where(want >= 0, want, 0).sum() / (2.0 * (want >= 0).sum())
        assert abs(obs - af[:, pop].mean()) < 0.01


def test_student_t_device_matches_oracle():
    from hail_b200 import _lib
    ctx = _lib.context(0)
    t = np.concatenate([np.linspace(-45, 45, 721), [0.0, 1e-12, -1e-9, 1.73, 1.74, np.nan, np.inf, -np.inf]])
    d_t = torch.from_numpy(t).cuda()
    for df in (1, 2, 5, 30, 994, 399989, 499989):
        d_p = torch.empty_like(d_t)
        d_l = torch.empty_like(d_t)
        ctx.check(ctx.lib.lrr_student_t_two_sided(ctx.handle, d_t.data_ptr(), t.size, float(df), d_p.data_ptr(),
                                                  d_l.data_ptr(), None))
        torch.cuda.synchronize()
        p, l10 = d_p.cpu().numpy(), d_l.cpu().numpy()
        want = O.two_sided_p(t, df)
        fin = np.isfinite(t)
        big = fin & (want > 1e-300)
        assert np.allclose(p[big], want[big], rtol=1e-9, atol=0), df
        assert np.isnan(p[np.isnan(t)]).all() and (p[np.isinf(t)] == 0).all()
        # log scale, including where p underflows (BASELINE: log-scale comparison below 1e-300)
        for i in np.nonzero(fin)[0][::40]:
            wl = O.log_two_sided_p(t[i], df) / np.log(10.0)
            assert abs(l10[i] - wl) <= 1e-8 * max(1.0, abs(wl)), (df, t[i], l10[i], wl)


@pytest.mark.parametrize("kernel", KERNELS)
def test_reference_golden_with_cov(kernel):  # TS:245-284
    hb = _hb()
    x, y, cov, doc = load_regression_linear()
    mt = _mt_from_dosage(x, pheno=y, Cov1=cov[:, 0], Cov2=cov[:, 1])
    ht = hb.linear_regression_rows(y=mt.pheno, x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.Cov1, mt.Cov2],
                                   _kernel=kernel)
    exp = doc["expected"]["with_cov"]
    assert (ht.n == 6).all()
    for pos in ("1", "2", "3"):
        for f, v in exp[pos].items():
            assert abs(ht[f][int(pos) - 1] - v) < 5e-7, (pos, f)
    for v in exp["nan_se"]:
        assert np.isnan(ht.standard_error[v - 1]) and np.isnan(ht.t_stat[v - 1]) and np.isnan(ht.p_value[v - 1])
    # against the oracle for the non-degenerate rows (rows 6-10 are roundoff garbage in the reference)
    want = O.linreg_group(x, y[:, None], np.column_stack([np.ones(8), cov]))
    got = _as_oracle_dict(ht)
    good = slice(0, 5)
    assert_fields_close({k: v[good] for k, v in got.items()}, {k: v[good] for k, v in want.items() if k != "_d"},
                        t_floor=1e-9)
    assert np.array_equal(got["sum_x"], want["sum_x"], equal_nan=True) or np.allclose(got["sum_x"], want["sum_x"], rtol=1e-15)


@pytest.mark.parametrize("kernel", KERNELS)
def test_reference_golden_no_covariates(kernel):  # TS:223-234
    hb = _hb()
    x, y, cov, doc = load_regression_linear()
    mt = _mt_from_dosage(x, pheno=y)
    with pytest.warns(UserWarning, match="no intercept"):
        ht = hb.linear_regression_rows(y=mt.pheno, x=mt.GT.n_alt_alleles(), covariates=[], _kernel=kernel)
    for f, v in doc["expected"]["no_intercept"]["1"].items():
        assert abs(ht[f][0] - v) < 5e-7, f


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("kind", ["is_case", "quant"])
def test_reference_golden_fam(kernel, kind):  # TS:366-424
    hb = _hb()
    x, y, cov, doc = load_regression_linear(fam_pheno=kind)
    mt = _mt_from_dosage(x, pheno=y, Cov1=cov[:, 0], Cov2=cov[:, 1])
    ht = hb.linear_regression_rows(y=mt.pheno, x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.Cov1, mt.Cov2],
                                   _kernel=kernel)
    exp = doc["expected"]["with_cov"]
    for pos in ("1", "2"):
        for f, v in exp[pos].items():
            assert abs(ht[f][int(pos) - 1] - v) < 5e-7, (pos, f)
    for v in (6, 7, 8, 9, 10):
        assert np.isnan(ht.standard_error[v - 1])


@pytest.mark.parametrize("kernel", KERNELS)
def test_linreg_basic_call_shapes_agree(kernel):  # TS:62-93
    hb = _hb()
    x, y, cov, doc = load_regression_linear(pheno_missing_zero=False)
    mt = _mt_from_dosage(x, pheno=y, cov={"Cov1": cov[:, 0], "Cov2": cov[:, 1]})
    mt = mt.annotate_entries(x=mt.GT.n_alt_alleles())
    covs = [1.0, mt.cov.Cov1, mt.cov.Cov2]
    t1 = hb.linear_regression_rows(y=mt.pheno, x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.cov.Cov1, mt.cov.Cov2 + 1 - 1], _kernel=kernel)
    t2 = hb.linear_regression_rows(y=mt.pheno, x=mt.x, covariates=covs, _kernel=kernel)
    t3 = hb.linear_regression_rows(y=[mt.pheno], x=mt.x, covariates=covs, _kernel=kernel)
    t4 = hb.linear_regression_rows(y=[mt.pheno, mt.pheno], x=mt.x, covariates=covs, _kernel=kernel)
    p1 = t1.select(p=t1.p_value)
    assert p1._same(t2.select(p=t2.p_value))
    assert p1._same(t3.select(p=t3.p_value[:, 0]))
    assert p1._same(t4.select(p=t4.p_value[:, 0]))
    assert p1._same(t4.select(p=t4.p_value[:, 1]))
    assert list(t1.row) == ["locus", "alleles", "n", "sum_x", "y_transpose_x", "beta", "standard_error", "t_stat", "p_value"]
    assert t1.n.dtype == np.int32 and t1.beta.ndim == 1 and t3.beta.shape == (10, 1) and t4.beta.shape == (10, 2)


@pytest.mark.parametrize("kernel", KERNELS)
def test_linreg_chained(kernel):  # TS:134-221
    hb = _hb()
    x, y, cov, doc = load_regression_linear(pheno_missing_zero=False)
    mt = _mt_from_dosage(x, pheno=y, cov={"Cov1": cov[:, 0], "Cov2": cov[:, 1]})
    mt = mt.annotate_entries(x=mt.GT.n_alt_alleles())

    def eq(a, b):
        a, b = np.asarray(a), np.asarray(b)
        return bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())

    t1 = hb.linear_regression_rows(y=[[mt.pheno], [mt.pheno]], x=mt.x, covariates=[1, mt.cov.Cov1, mt.cov.Cov2], _kernel=kernel)
    assert t1.n.shape == (10, 2) and eq(t1.n[:, 0], t1.n[:, 1]) and eq(t1.sum_x[:, 0], t1.sum_x[:, 1])
    for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
        assert eq(t1[f][0], t1[f][1]), f

    mt2 = mt.filter_cols(mt.cov.Cov2 >= 0)
    mt3 = mt.filter_cols(mt.cov.Cov2 <= 0)
    t2 = hb.linear_regression_rows(y=mt2.pheno, x=mt2.x, covariates=[1, mt2.cov.Cov1], _kernel=kernel)
    t3 = hb.linear_regression_rows(y=mt3.pheno, x=mt3.x, covariates=[1, mt3.cov.Cov1], _kernel=kernel)
    chained = hb.linear_regression_rows(
        y=[[mt.pheno.or_missing_unless(mt.cov.Cov2 >= 0)], [mt.pheno.or_missing_unless(mt.cov.Cov2 <= 0)]],
        x=mt.x, covariates=[1, mt.cov.Cov1], _kernel=kernel)
    for g, sep in enumerate((t2, t3)):
        assert eq(chained.n[:, g], sep.n) and eq(chained.sum_x[:, g], sep.sum_x)
        for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
            assert eq(chained[f][g][:, 0], sep[f]), (g, f)

    phenos = [mt.pheno.or_missing_unless(mt.cov.Cov2 >= -1), mt.pheno.or_missing_unless(mt.cov.Cov2 <= 1)]
    t4 = hb.linear_regression_rows(phenos, mt.x, covariates=[1], _kernel=kernel)
    t5 = hb.linear_regression_rows([phenos], mt.x, covariates=[1], _kernel=kernel)
    assert eq(t4.n, t5.n[:, 0]) and eq(t4.sum_x, t5.sum_x[:, 0])
    for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
        assert hb.Table({"v": t4[f]}, n_rows=10)._same(hb.Table({"v": t5[f][0]}, n_rows=10)), f


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("P,K,intercept", [(1, 2, True), (3, 4, True), (2, 3, False), (1, 0, False), (5, 1, True)])
def test_fastlmm_parity_vs_oracle(kernel, P, K, intercept):
    hb = _hb()
    z = np.load(os.path.join(GOLDEN, "fastlmm.npz"))
    N, M = int(z["n_samples"]), int(z["n_variants"])
    rows = obed.bed_body(z["bed"], N, M)
    rng = np.random.default_rng(10 * P + K)
    ys = np.column_stack([z["pheno"]] + [rng.normal(size=N) for _ in range(P - 1)])
    ys[rng.random(ys.shape) < 0.03] = np.nan
    cols = ([np.ones(N)] if intercept else []) + [z["cov"][:, 0]] + [rng.normal(size=N) + 0.5 for _ in range(8)]
    cov = np.column_stack(cols)[:, :K] if K else np.empty((N, 0))
    x = obed.decode_rows(rows, N)
    want = O.linreg_group(x, ys, cov)
    gt = hb.PackedGenotypes.from_bed_rows(rows, N)
    mt = hb.MatrixTable(gt, cols={f"y{i}": ys[:, i] for i in range(P)} | {f"c{i}": cov[:, i] for i in range(K)})
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ht = hb.linear_regression_rows(y=[mt[f"y{i}"] for i in range(P)], x=mt.GT.n_alt_alleles(),
                                       covariates=[mt[f"c{i}"] for i in range(K)], _kernel=kernel)
    got = _as_oracle_dict(ht)
    nondeg = np.isfinite(want["standard_error"]).all(axis=1)
    assert nondeg.sum() > 0.9 * M
    assert_fields_close({k: v[nondeg] for k, v in got.items()}, {k: v[nondeg] for k, v in want.items() if k != "_d"},
                        t_floor=1e-9, ctx=f"{kernel} P={P} K={K}")
    # bit-exact integer outputs: n, missing counts, and sum_x on missing-free rows
    idx = O.complete_samples(ys, cov)[2]
    nm = np.isnan(x[:, idx]).sum(axis=1)
    assert np.array_equal(ht.n_missing, nm)
    clean = nm == 0
    assert np.array_equal(got["sum_x"][clean], want["sum_x"][clean])


@pytest.mark.parametrize("kernel", KERNELS)
def test_bn_missing_chained_vs_oracle(kernel):
    """BASELINE config 3 in miniature: 25 % missing calls, y=[[y1],[y2]] with per-group phenotype missingness."""
    hb = _hb()
    N, M = 3000, 512
    mt = hb.balding_nichols_model(3, N, M, missing_rate=0.25, seed=3)
    rng = np.random.default_rng(5)
    dos = mt.genotypes.to_dosage().astype(np.float64)
    dos[dos < 0] = np.nan
    cov = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(9)])
    y1 = rng.normal(size=N) + 0.3 * np.nan_to_num(dos[7]) ; y2 = rng.normal(size=N)
    y1[rng.random(N) < 0.1] = np.nan
    y2[rng.random(N) < 0.2] = np.nan
    mt = mt.annotate_cols(y1=y1, y2=y2, **{f"c{i}": cov[:, i] for i in range(10)})
    ht = hb.linear_regression_rows(y=[[mt.y1], [mt.y2]], x=mt.GT.n_alt_alleles(),
                                   covariates=[mt[f"c{i}"] for i in range(10)], _kernel=kernel)
    want = O.linreg_chained(dos, [y1[:, None], y2[:, None]], cov)
    for g in range(2):
        got = {"n": ht.n[:, g], "sum_x": ht.sum_x[:, g]}
        for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
            got[f] = ht[f][g]
        assert_fields_close(got, {k: v for k, v in want[g].items() if k != "_d"}, t_floor=1e-9, ctx=f"{kernel} g={g}")
        idx = O.complete_samples([y1, y2][g][:, None], cov)[2]
        assert np.array_equal(ht.n_missing[g], np.isnan(dos[:, idx]).sum(axis=1))


@pytest.mark.parametrize("kernel", KERNELS)
def test_edge_cases(kernel):
    hb = _hb()
    rng = np.random.default_rng(2)
    N = 70
    x = rng.integers(0, 3, size=(9, N)).astype(np.float64)
    x[0] = np.nan            # all-missing variant -> every statistic NaN, sum_x NaN (RU:52)
    x[1] = 1.0               # constant -> degenerate
    x[2, ::2] = np.nan
    y = rng.normal(size=N)
    mt = _mt_from_dosage(x, y=y, c=rng.normal(size=N))
    ht = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.c], _kernel=kernel)
    assert np.isnan(ht.sum_x[0]) and np.isnan(ht.beta[0]) and np.isnan(ht.p_value[0])
    assert ht.sum_x[1] == N and np.isnan(ht.standard_error[1])
    want = O.linreg_group(x, y[:, None], np.column_stack([np.ones(N), mt.c.values]))
    got = _as_oracle_dict(ht)
    assert_fields_close({k: v[2:] for k, v in got.items()}, {k: v[2:] for k, v in want.items() if k != "_d"}, t_floor=1e-9)
    # empty row range
    empty = hb.MatrixTable(hb.PackedGenotypes.empty(0, N), cols={"y": y})
    ht0 = hb.linear_regression_rows(y=empty.y, x=empty.GT.n_alt_alleles(), covariates=[1.0], _kernel=kernel)
    assert ht0.count() == 0 and ht0.beta.shape == (0,)


def test_fatal_conditions_match_reference_messages():
    hb = _hb()
    x, y, cov, doc = load_regression_linear()
    mt = _mt_from_dosage(x, pheno=y, Cov1=cov[:, 0], Cov2=cov[:, 1], allmiss=np.full(8, np.nan))
    with pytest.raises(hb.FatalError, match="degrees of freedom"):
        hb.linear_regression_rows(mt.pheno, mt.GT.n_alt_alleles(),
                                  [1.0, mt.Cov1, mt.Cov2, mt.Cov1 * mt.Cov1, mt.Cov2 * mt.Cov2, mt.Cov1 * mt.Cov2])
    with pytest.raises(hb.FatalError, match="No complete samples"):
        hb.linear_regression_rows(mt.allmiss, mt.GT.n_alt_alleles(), [1.0])


# ---------------------------------------------------------------------------------------------
# BASELINE.json sample sizes (400k / 500k samples): parity against the C oracle on a variant slice, plus
# size-independent properties (kernel-vs-kernel agreement, exact integer fields, planted signal ranking).
# ---------------------------------------------------------------------------------------------
def _big_case(N, M, K, missing_rate, seed):
    hb = _hb()
    from hail_b200 import bn
    rng = np.random.Generator(np.random.Philox(key=[seed, 77]))
    pop, th, _ = bn.bn_parameters(3, N, M, missing_rate=missing_rate, seed=seed)
    gt = bn.bn_fill(hb.PackedGenotypes.empty(M, N), pop, th, seed=seed)
    cov = np.column_stack([np.ones(N)] + [rng.standard_normal(N) for _ in range(K - 1)])
    # pull the bed bytes of the same store back for the CPU side
    from hail_b200 import _lib
    ctx = _lib.context(0)
    bed_stride = (N + 3) // 4
    d_bed = torch.empty((M, bed_stride), dtype=torch.uint8, device="cuda")
    ctx.check(ctx.lib.lrr_unpack_bed(ctx.handle, gt.data.data_ptr(), gt.stride, M, N, d_bed.data_ptr(), bed_stride, None))
    torch.cuda.synchronize()
    return gt, d_bed.cpu().numpy(), cov, rng


@pytest.mark.parametrize("N,missing_rate", [(400_000, 0.0), (400_000, 0.25), (500_000, 0.0)])
def test_full_sample_size_parity_vs_c_oracle(N, missing_rate):
    """C2 / C3 / C5 sample counts, K = 10 (intercept + 9 PCs); variant slice sized for seconds of CPU."""
    hb = _hb()
    from oracle import c_oracle
    M, K = 1024, 10
    gt, bed_rows, cov, rng = _big_case(N, M, K, missing_rate, seed=21)
    dos0 = obed.decode_rows(bed_rows[:1], N)[0]
    y = rng.standard_normal(N) + 0.02 * np.nan_to_num(dos0)      # variant 0 is causal
    y[rng.random(N) < 0.05] = np.nan
    want = c_oracle.linreg_group_bed(bed_rows, N, y[:, None], cov)
    mt = hb.MatrixTable(gt, cols={"y": y, **{f"c{i}": cov[:, i] for i in range(1, K)}})
    covs = [1.0] + [mt[f"c{i}"] for i in range(1, K)]
    res = {}
    for kernel in KERNELS:
        ht = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=covs, _kernel=kernel, _log10_p=True)
        got = _as_oracle_dict(ht)
        assert_fields_close(got, {k: v for k, v in want.items() if k != "_d"}, t_floor=1e-9,
                            ctx=f"{kernel} N={N} miss={missing_rate}")
        res[kernel] = ht
        if kernel == "tc4" and missing_rate == 0.0:
            # BASELINE's headline shape runs 80 MMA columns: 9 six-digit covariate columns + 13 + 11 digits for the phenotype
            # and its fitted-value column + the "ones" row (Gaussian columns must not cost a tail digit: DESIGN.md 5.1c)
            from hail_b200 import _lib
            assert _lib.context(0).last_sweep_shape == (1, 80, 80, 79), _lib.context(0).last_sweep_shape
        # the causal variant is by far the strongest signal and its log10 p is finite
        assert int(np.nanargmin(ht.p_value)) == 0 and np.isfinite(ht.log10_p[0]) and ht.log10_p[0] < -6
    # the kernels agree far below the tolerance (exact integer paths vs float64 FMA path)
    for name in TC_KERNELS:
        a, b = res[name], res["fp64"]
        assert np.array_equal(a.n, b.n) and np.array_equal(a.n_missing, b.n_missing)
        clean = a.n_missing == 0
        assert np.array_equal(a.sum_x[clean], b.sum_x[clean])
        assert np.nanmax(np.abs(a.t_stat - b.t_stat)) < 1e-8
    # integer outputs of the two tensor-core sweeps are identical bit for bit, missing calls or not
    assert np.array_equal(res["tc"].sum_x, res["tc4"].sum_x)


@pytest.mark.parametrize("tck", TC_KERNELS)
def test_tc_kernel_is_deterministic_and_order_independent(tck):
    """Exact integer accumulation: the same rows give bit-identical results wherever they sit in the sweep."""
    hb = _hb()
    N, M = 100_000, 1500
    gt, _, cov, rng = _big_case(N, M, 4, 0.1, seed=5)
    y = rng.standard_normal(N)
    cols = {"y": y, **{f"c{i}": cov[:, i] for i in range(1, 4)}}
    mt = hb.MatrixTable(gt, cols=cols)
    covs = [1.0] + [mt[f"c{i}"] for i in range(1, 4)]
    full = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=covs, _kernel=tck)
    again = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=covs, _kernel=tck)
    sub = hb.MatrixTable(gt.rows(512, 1400), cols=cols)   # different tile alignment / CTA assignment
    part = hb.linear_regression_rows(y=sub.y, x=sub.GT.n_alt_alleles(),
                                     covariates=[1.0] + [sub[f"c{i}"] for i in range(1, 4)], _kernel=tck)
    for f in ("sum_x", "y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
        assert np.array_equal(full[f], again[f], equal_nan=True), f
        assert np.array_equal(full[f][512:1400], part[f], equal_nan=True), f


def test_mixed_one_plane_two_plane_tiles_many_tiles_per_cta():
    """Tiles with and without missing calls interleaved, > 2 tiles per CTA, odd tail: exercises the mode switches
    of the TMEM ring and the cluster padding tiles."""
    hb = _hb()
    rng = np.random.default_rng(9)
    N, M = 1500, 148 * 128 * 2 + 128 * 5 + 37
    x = rng.integers(0, 3, size=(M, N)).astype(np.int8)
    tile = np.arange(M) // 128
    miss_rows = (tile % 3 == 1) | (tile % 7 == 0)
    mask = (rng.random((M, N)) < 0.08) & miss_rows[:, None]
    x[mask] = -1
    gt = hb.PackedGenotypes.from_dosage(x)
    assert np.array_equal(gt.row_flags.cpu().numpy().astype(bool), (x < 0).any(axis=1))
    cov = np.column_stack([np.ones(N), rng.normal(size=(N, 2))])
    y = rng.normal(size=N)
    mt = hb.MatrixTable(gt, cols={"y": y, "c1": cov[:, 1], "c2": cov[:, 2]})
    xf = np.where(x < 0, np.nan, x).astype(np.float64)
    want = O.linreg_group(xf, y[:, None], cov)
    for kernel in KERNELS:
        ht = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.c1, mt.c2], _kernel=kernel)
        got = _as_oracle_dict(ht)
        assert_fields_close(got, {k: v for k, v in want.items() if k != "_d"}, t_floor=1e-9, ctx=kernel)
        assert np.array_equal(ht.n_missing, (x < 0).sum(axis=1))
    # unknown flags (None) must give the same answer: every tile is then treated as possibly missing
    gt2 = hb.PackedGenotypes(gt.data, M, N, None)
    mt2 = hb.MatrixTable(gt2, cols={"y": y, "c1": cov[:, 1], "c2": cov[:, 2]})
    for tck in TC_KERNELS:
        h1 = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.c1, mt.c2], _kernel=tck)
        h2 = hb.linear_regression_rows(y=mt2.y, x=mt2.GT.n_alt_alleles(), covariates=[1.0, mt2.c1, mt2.c2], _kernel=tck)
        assert np.array_equal(h2.beta, h1.beta, equal_nan=True) and np.array_equal(h2.p_value, h1.p_value, equal_nan=True)


@pytest.mark.parametrize("kernel", KERNELS)
def test_many_phenotypes_multi_pass(kernel):
    """BASELINE config 4 in miniature: many phenotypes -> more digit columns than one sweep holds (multi-pass)."""
    hb = _hb()
    N, M, P, K = 3000, 640, 45, 6
    mt = hb.balding_nichols_model(3, N, M, missing_rate=0.02, seed=13)
    rng = np.random.default_rng(14)
    dos = mt.genotypes.to_dosage().astype(np.float64)
    dos[dos < 0] = np.nan
    cov = np.column_stack([np.ones(N)] + [rng.normal(size=N) for _ in range(K - 1)])
    ys = rng.normal(size=(N, P)) + 0.1 * np.nan_to_num(dos[:P].T)
    ys[rng.random(N) < 0.03, 0] = np.nan
    mt = mt.annotate_cols(**{f"y{i}": ys[:, i] for i in range(P)}, **{f"c{i}": cov[:, i] for i in range(1, K)})
    ht = hb.linear_regression_rows(y=[mt[f"y{i}"] for i in range(P)], x=mt.GT.n_alt_alleles(),
                                   covariates=[1.0] + [mt[f"c{i}"] for i in range(1, K)], _kernel=kernel)
    want = O.linreg_group(dos, ys, cov)
    assert ht.beta.shape == (M, P)
    assert_fields_close(_as_oracle_dict(ht), {k: v for k, v in want.items() if k != "_d"}, t_floor=1e-9, ctx=kernel)


def test_multi_pass_with_sparse_missing_tiles():
    """Several wide passes over data where only SOME tile pairs hold missing calls, > 2 tiles per CTA: the split
    plane-c / plane-m sweeps (raw sums left by the first, rows finished by the second, untouched pairs skipped) must
    give the rows of the oracle, the exact counts, and the same rows as the int8 kernel's two-plane passes."""
    hb = _hb()
    rng = np.random.default_rng(21)
    N, M, P, K = 1500, 148 * 128 * 2 + 128 * 5 + 37, 20, 4
    x = rng.integers(0, 3, size=(M, N)).astype(np.int8)
    pair = np.arange(M) // 256
    miss_rows = (pair % 3 == 1) | (pair % 11 == 0)
    miss_rows[-20:] = True                       # the ragged last tile too
    mask = (rng.random((M, N)) < 0.05) & miss_rows[:, None]
    x[mask] = -1
    x[300] = -1                                   # an all-missing variant inside a flagged pair
    gt = hb.PackedGenotypes.from_dosage(x)
    cov = np.column_stack([np.ones(N), rng.normal(size=(N, K - 1))])
    ys = rng.normal(size=(N, P))
    cols = {f"y{i}": ys[:, i] for i in range(P)}
    cols.update({f"c{i}": cov[:, i] for i in range(1, K)})
    mt = hb.MatrixTable(gt, cols=cols)
    xf = np.where(x < 0, np.nan, x).astype(np.float64)
    want = O.linreg_group(xf, ys, cov)
    floor = ytx_floor_of(xf, cov, want["standard_error"])
    rows = {}
    for kernel in ("tc4", "auto", "tc"):
        ht = hb.linear_regression_rows(y=[mt[f"y{i}"] for i in range(P)], x=mt.GT.n_alt_alleles(),
                                       covariates=[1.0] + [mt[f"c{i}"] for i in range(1, K)], _kernel=kernel)
        assert np.array_equal(ht.n_missing, (x < 0).sum(axis=1)), kernel
        # 770k dot products: a few cancel to ~1e-5 of their scale; P > 2 is the many-phenotype profile (its stated floor)
        assert_fields_close(_as_oracle_dict(ht), {k: v for k, v in want.items() if k != "_d"}, t_floor=1e-9, ctx=kernel,
                            ytx_abs=None if kernel == "tc" else floor)
        rows[kernel] = ht
    assert np.array_equal(rows["tc4"].sum_x, rows["tc"].sum_x, equal_nan=True)
    # chained groups on the same data (different complete-sample sets per group), still more columns than one pass
    y2 = ys.copy()
    y2[rng.random(N) < 0.1] = np.nan
    mt2 = mt.annotate_cols(**{f"z{i}": y2[:, i] for i in range(P)})
    ht = hb.linear_regression_rows(y=[[mt2[f"y{i}"] for i in range(P)], [mt2[f"z{i}"] for i in range(P)]],
                                   x=mt2.GT.n_alt_alleles(), covariates=[1.0] + [mt2[f"c{i}"] for i in range(1, K)])
    wantc = O.linreg_chained(xf, [ys, y2], cov)
    for g in range(2):
        got = {"n": ht.n[:, g], "sum_x": ht.sum_x[:, g]}
        for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
            got[f] = ht[f][g]
        idx = O.complete_samples([ys, y2][g], cov)[2]
        assert_fields_close(got, {k: v for k, v in wantc[g].items() if k != "_d"}, t_floor=1e-9, ctx=f"chained g={g}",
                            ytx_abs=ytx_floor_of(xf[:, idx], cov[idx], wantc[g]["standard_error"]))
        assert np.array_equal(ht.n_missing[g], (x[:, idx] < 0).sum(axis=1))


@pytest.mark.parametrize("kernel", KERNELS)
def test_many_chained_groups(kernel):
    """More groups than one sweep's segment table holds (5 chained groups with different missingness)."""
    hb = _hb()
    N, M, G = 1200, 300, 5
    mt = hb.balding_nichols_model(3, N, M, missing_rate=0.1, seed=17)
    rng = np.random.default_rng(18)
    dos = mt.genotypes.to_dosage().astype(np.float64)
    dos[dos < 0] = np.nan
    cov = np.column_stack([np.ones(N), rng.normal(size=(N, 2))])
    ys = []
    for g in range(G):
        y = rng.normal(size=(N, 1 + g % 2))
        y[rng.random(N) < 0.05 * (g + 1)] = np.nan
        ys.append(y)
    cols = {f"c{i}": cov[:, i] for i in range(1, 3)}
    for g in range(G):
        for j in range(ys[g].shape[1]):
            cols[f"y{g}_{j}"] = ys[g][:, j]
    mt = mt.annotate_cols(**cols)
    ht = hb.linear_regression_rows(y=[[mt[f"y{g}_{j}"] for j in range(ys[g].shape[1])] for g in range(G)],
                                   x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.c1, mt.c2], _kernel=kernel)
    want = O.linreg_chained(dos, ys, cov)
    for g in range(G):
        got = {"n": ht.n[:, g], "sum_x": ht.sum_x[:, g]}
        for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
            got[f] = ht[f][g]
        assert_fields_close(got, {k: v for k, v in want[g].items() if k != "_d"}, t_floor=1e-9, ctx=f"{kernel} g={g}")


def test_prepared_buffers_are_reused_across_calls_and_returned_by_trim():
    """The 4-bit sweep's prepared buffers live in grow-only pools (no cudaFree on the per-call path): calls whose groups
    shrink, grow and change shape in turn give the float64 kernel's rows every time, lrr_trim hands the pools back once no
    groups are held, and the next call simply builds them again."""
    hb = _hb()
    from hail_b200 import _lib
    ctx = _lib.context(0)
    rng = np.random.default_rng(44)
    N, M = 3000, 700
    mt = hb.balding_nichols_model(3, N, M, missing_rate=0.02, seed=9)
    cov = rng.normal(size=(N, 12))
    ys = rng.normal(size=(N, 6))
    mt = mt.annotate_cols(**{f"c{i}": cov[:, i] for i in range(12)}, **{f"y{i}": ys[:, i] for i in range(6)})

    def both(y, K):
        covs = [1.0] + [mt[f"c{i}"] for i in range(K)]
        a = hb.linear_regression_rows(y=y, x=mt.GT.n_alt_alleles(), covariates=covs, _kernel="tc4")
        b = hb.linear_regression_rows(y=y, x=mt.GT.n_alt_alleles(), covariates=covs, _kernel="fp64")
        assert np.array_equal(a.n, b.n) and np.array_equal(a.sum_x, b.sum_x)
        for f in ("beta", "standard_error", "t_stat"):
            fa, fb = a[f], b[f]
            pairs = [(fa, fb)] if isinstance(fa, np.ndarray) else list(zip(fa, fb))   # chained: one array per group
            for ga, gb in pairs:
                assert O.d_eq(np.asarray(ga), np.asarray(gb), 1e-6).all(), f
        return a

    first = both(mt.y0, 3)
    both([mt.y0, mt.y1, mt.y2, mt.y3, mt.y4, mt.y5], 12)      # many more digit columns: every pool grows
    both([[mt.y0], [mt.y1, mt.y2]], 5)                        # chained groups, fewer columns: the pools are reused
    again = both(mt.y0, 3)
    assert np.array_equal(first.beta, again.beta, equal_nan=True) and np.array_equal(first.p_value, again.p_value, equal_nan=True)
    ctx.check(ctx.lib.lrr_clear_groups(ctx.handle))
    ctx.check(ctx.lib.lrr_trim(ctx.handle))
    after = both(mt.y0, 3)
    assert np.array_equal(first.beta, after.beta, equal_nan=True)


# ---------------------------------------------------------------------------------------------
# host-resident input: the streaming loop (lrr_stream_*) must give exactly the resident path's rows
@pytest.mark.parametrize("block,depth", [(0, 0), (128, 2), (300, 3), (1000, 1), (4096, 5)])
def test_stream_from_host_bed_matches_resident(block, depth):
    hb = _hb()
    z = np.load(os.path.join(GOLDEN, "fastlmm.npz"))
    N, M = int(z["n_samples"]), int(z["n_variants"])
    rows = obed.bed_body(z["bed"], N, M)
    rng = np.random.default_rng(3)
    ys = np.column_stack([z["pheno"], rng.normal(size=N)])
    ys[rng.random(ys.shape) < 0.03] = np.nan
    cov = np.column_stack([np.ones(N), z["cov"][:, 0], rng.normal(size=N)])
    cols = {"y0": ys[:, 0], "y1": ys[:, 1], "c1": cov[:, 1], "c2": cov[:, 2]}
    host = hb.HostBedGenotypes(rows, N)
    assert host.rows.is_pinned() and host.n_variants == M
    mt_h = hb.MatrixTable(host, cols=cols)
    mt_d = hb.MatrixTable(hb.PackedGenotypes.from_bed_rows(rows, N), cols=cols)
    for y_of in (lambda mt: [mt.y0, mt.y1], lambda mt: [[mt.y0], [mt.y1]]):   # Single and Chained
        th = hb.linear_regression_rows(y=y_of(mt_h), x=mt_h.GT.n_alt_alleles(), covariates=[1.0, mt_h.c1, mt_h.c2],
                                       _kernel="tc", _stream_block=block, _stream_depth=depth)
        td = hb.linear_regression_rows(y=y_of(mt_d), x=mt_d.GT.n_alt_alleles(), covariates=[1.0, mt_d.c1, mt_d.c2],
                                       _kernel="tc")
        assert np.array_equal(th.n, td.n) and np.array_equal(np.asarray(th.n_missing), np.asarray(td.n_missing))
        for f in ("sum_x", "y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
            a, b = th[f], td[f]
            if isinstance(a, (list, tuple)) or hasattr(a, "__iter__") and not isinstance(a, np.ndarray):
                for ga, gb in zip(a, b):
                    assert np.array_equal(np.asarray(ga), np.asarray(gb), equal_nan=True), f
            else:
                assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True), f   # same kernels, same bits
    # and against the oracle
    want = O.linreg_group(obed.decode_rows(rows, N), ys, cov)
    th = hb.linear_regression_rows(y=[mt_h.y0, mt_h.y1], x=mt_h.GT.n_alt_alleles(), covariates=[1.0, mt_h.c1, mt_h.c2],
                                   _stream_block=block, _stream_depth=depth)
    got = _as_oracle_dict(th)
    nondeg = np.isfinite(want["standard_error"]).all(axis=1)
    assert_fields_close({k: v[nondeg] for k, v in got.items()}, {k: v[nondeg] for k, v in want.items() if k != "_d"},
                        t_floor=1e-9, ctx=f"stream block={block} depth={depth}")


def test_stream_empty_errors_and_file(tmp_path):
    hb = _hb()
    from hail_b200 import _lib
    z = np.load(os.path.join(GOLDEN, "fastlmm.npz"))
    N, M = int(z["n_samples"]), int(z["n_variants"])
    rows = obed.bed_body(z["bed"], N, M)
    # file -> page-locked memory
    path = tmp_path / "t.bed"
    path.write_bytes(bytes([0x6C, 0x1B, 0x01]) + rows.tobytes())
    host = hb.HostBedGenotypes.from_bed_file(str(path), N, M)
    assert np.array_equal(host.rows.numpy(), rows)
    with pytest.raises(ValueError):
        hb.HostBedGenotypes.from_bed_file(str(path), N, M + 1)
    # a FatalError in the prologue (no degrees of freedom) must close the stream so the next call works
    mt = hb.MatrixTable(host, cols={"y": z["pheno"], **{f"c{i}": np.random.default_rng(i).normal(size=N) for i in range(260)}})
    with pytest.raises(hb.FatalError):
        hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0] + [mt[f"c{i}"] for i in range(260)])
    t = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0])
    assert t.count() == M and np.isfinite(np.asarray(t.beta)).sum() > 0.9 * M
    # zero rows
    e = hb.MatrixTable(hb.HostBedGenotypes(rows[:0], N), cols={"y": z["pheno"]})
    t0 = hb.linear_regression_rows(y=e.y, x=e.GT.n_alt_alleles(), covariates=[1.0])
    assert t0.count() == 0
    # two open streams on one context are refused
    ctx = _lib.context(0)
    import ctypes
    h1, h2 = ctypes.c_void_p(), ctypes.c_void_p()
    ctx.check(ctx.lib.lrr_stream_begin(ctx.handle, ctypes.byref(h1), host.rows.data_ptr(), M, host.bed_stride, N, 0, 0))
    rc = ctx.lib.lrr_stream_begin(ctx.handle, ctypes.byref(h2), host.rows.data_ptr(), M, host.bed_stride, N, 0, 0)
    assert rc != 0 and b"still open" in ctx.lib.lrr_last_error(ctx.handle)
    ctx.lib.lrr_stream_end(ctx.handle, h1)


# ---------------------------------------------------------------------------------------------
# weights (WLS): the reference routes it through _linear_regression_rows_nd (statgen.py:344-345, 557-581, 636-660)
def test_weighted_linear_regression_vs_oracle():
    hb = _hb()
    z = np.load(os.path.join(GOLDEN, "fastlmm.npz"))
    N, M = int(z["n_samples"]), int(z["n_variants"])
    rows = obed.bed_body(z["bed"], N, M)
    x = obed.decode_rows(rows, N)
    rng = np.random.default_rng(31)
    ys = np.column_stack([z["pheno"], rng.normal(size=N)])
    ys[rng.random(ys.shape) < 0.03] = np.nan
    cov = np.column_stack([np.ones(N), z["cov"][:, 0], rng.normal(size=N)])
    w1 = rng.uniform(0.2, 4.0, size=N)
    w1[rng.random(N) < 0.04] = np.nan                      # missing weights drop the sample (test_statgen.py:610-660)
    w2 = np.arange(N, dtype=np.float64) + 5.0             # test_statgen.py:596 uses col_idx + 5
    mt = hb.MatrixTable(hb.PackedGenotypes.from_bed_rows(rows, N),
                        cols={"y0": ys[:, 0], "y1": ys[:, 1], "c1": cov[:, 1], "c2": cov[:, 2], "w1": w1, "w2": w2})
    covs = [1.0, mt.c1, mt.c2]
    # one weight expression, list of phenotypes
    ht = hb.linear_regression_rows(y=[mt.y0, mt.y1], x=mt.GT.n_alt_alleles(), covariates=covs, weights=mt.w1)
    want = O.linreg_group_weighted(x, ys, cov, w1)
    nondeg = np.isfinite(want["standard_error"]).all(axis=1)
    assert nondeg.sum() > 0.9 * M
    assert_fields_close({k: v[nondeg] for k, v in _as_oracle_dict(ht).items()},
                        {k: v[nondeg] for k, v in want.items() if k != "_d"}, t_floor=1e-9, ctx="weighted")
    # chained: one weight per group (test_statgen.py:595-607)
    hc = hb.linear_regression_rows(y=[[mt.y0], [mt.y1]], x=mt.GT.n_alt_alleles(), covariates=covs, weights=[mt.w1, mt.w2])
    for g, (yy, ww) in enumerate(((ys[:, :1], w1), (ys[:, 1:], w2))):
        wg = O.linreg_group_weighted(x, yy, cov, ww)
        got = {"n": hc.n[:, g], "sum_x": hc.sum_x[:, g]}
        for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
            got[f] = hc[f][g]
        nd = np.isfinite(wg["standard_error"]).all(axis=1)
        assert_fields_close({k: v[nd] for k, v in got.items()}, {k: v[nd] for k, v in wg.items() if k != "_d"},
                            t_floor=1e-9, ctx=f"weighted chained g={g}")
    # unit weights reproduce the unweighted regression (statistics; sum_x is then the plain column sum)
    mt1 = mt.annotate_cols(one=np.ones(N))
    hu = hb.linear_regression_rows(y=mt1.y0, x=mt1.GT.n_alt_alleles(), covariates=[1.0, mt1.c1, mt1.c2], weights=mt1.one)
    h0 = hb.linear_regression_rows(y=mt1.y0, x=mt1.GT.n_alt_alleles(), covariates=[1.0, mt1.c1, mt1.c2])
    ok = np.isfinite(h0.standard_error)
    assert np.allclose(hu.beta[ok], h0.beta[ok], rtol=1e-7) and np.allclose(hu.p_value[ok], h0.p_value[ok], rtol=1e-6)
    # the tensor-core kernels refuse weighted groups instead of computing something else
    from hail_b200 import _lib
    with pytest.raises(_lib.LrrError, match="float64 kernel"):
        hb.linear_regression_rows(y=mt.y0, x=mt.GT.n_alt_alleles(), covariates=covs, weights=mt.w1, _kernel="tc4")


# ---------------------------------------------------------------------------------------------
# logistic regression, score test: R-derived golden values and oracle parity
def test_logistic_score_reference_golden():   # test_statgen.py:987-1021
    hb = _hb()
    from tests.test_oracle_logistic import load_regression_logistic
    doc, x, y, cov = load_regression_logistic()
    mt = _mt_from_dosage(x, y=y, c1=cov[:, 1], c2=cov[:, 2])
    ht = hb.logistic_regression_rows(test="score", y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.c1, mt.c2])
    for pos, want in doc["expected_score"].items():
        if pos == "constant":
            continue
        i = int(pos) - 1
        assert abs(ht.chi_sq_stat[i] - want["chi_sq_stat"]) < 5e-7 and abs(ht.p_value[i] - want["p_value"]) < 5e-7, pos
    for pos in doc["expected_score"]["constant"]:
        c = ht.chi_sq_stat[pos - 1]
        assert np.isnan(c) or c < 1e-6


def test_logistic_score_vs_oracle_and_errors():
    hb = _hb()
    from oracle import logreg_oracle as L
    z = np.load(os.path.join(GOLDEN, "fastlmm.npz"))
    N, M = int(z["n_samples"]), int(z["n_variants"])
    rows = obed.bed_body(z["bed"], N, M)
    x = obed.decode_rows(rows, N)
    rng = np.random.default_rng(41)
    cov = np.column_stack([np.ones(N), z["cov"][:, 0], rng.normal(size=N)])
    eta = 0.4 * cov[:, 1] - 0.3 * cov[:, 2] + 0.8 * np.nan_to_num(x[5]) - 0.5
    y1 = (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(np.float64)
    y2 = (rng.random(N) < 0.3).astype(np.float64)
    y1[rng.random(N) < 0.03] = np.nan
    mt = hb.MatrixTable(hb.PackedGenotypes.from_bed_rows(rows, N), rows={"rsid": np.arange(M)},
                        cols={"y1": y1, "y2": y2, "c1": cov[:, 1], "c2": cov[:, 2]})
    ht = hb.logistic_regression_rows("score", [mt.y1, mt.y2], mt.GT.n_alt_alleles(), [1.0, mt.c1, mt.c2],
                                     pass_through=["rsid"])
    assert ht.chi_sq_stat.shape == (M, 2) and list(ht.rsid) == list(range(M))
    # the Scala path keeps the samples complete for ALL phenotypes (LogisticRegression.scala:41-42)
    keep = ~np.isnan(y1)
    for col, yy in enumerate((y1, y2)):
        want = L.logreg_score(x[:, keep], yy[keep], cov[keep])
        ok = np.isfinite(want["chi_sq_stat"])
        assert ok.sum() > 0.9 * M
        assert np.allclose(ht.chi_sq_stat[ok, col], want["chi_sq_stat"][ok], rtol=1e-6, atol=1e-9), col
        assert np.allclose(ht.p_value[ok, col], want["p_value"][ok], rtol=1e-5, atol=1e-300), col
    assert int(np.nanargmin(ht.p_value[:, 0])) == 5        # the causal variant
    single = hb.logistic_regression_rows("score", mt.y2, mt.GT.n_alt_alleles(), [1.0, mt.c1, mt.c2])
    assert single.chi_sq_stat.shape == (M,)
    # errors: statgen.py:976-984, LogisticRegression.scala:44-61, 83-90
    with pytest.raises(ValueError, match="at least one covariate"):
        hb.logistic_regression_rows("score", mt.y2, mt.GT.n_alt_alleles(), [])
    with pytest.raises(ValueError, match="found no values for 'y'"):
        hb.logistic_regression_rows("score", [], mt.GT.n_alt_alleles(), [1.0])
    with pytest.raises(TypeError):
        hb.logistic_regression_rows("rao", mt.y2, mt.GT.n_alt_alleles(), [1.0])
    bad = mt.annotate_cols(q=rng.normal(size=N), one=np.ones(N), sep=np.where(y2 > 0, 9.0, -9.0))
    with pytest.raises(hb.FatalError, match="equal to 0 or 1"):
        hb.logistic_regression_rows("score", bad.q, bad.GT.n_alt_alleles(), [1.0])
    with pytest.raises(hb.FatalError, match="must be non-constant"):
        hb.logistic_regression_rows("score", bad.one, bad.GT.n_alt_alleles(), [1.0])
    with pytest.raises(hb.FatalError, match="Failed to fit logistic regression null model"):
        hb.logistic_regression_rows("score", bad.y2, bad.GT.n_alt_alleles(), [1.0, bad.sep])
    # the context is usable for the linear path afterwards
    lin = hb.linear_regression_rows(mt.y2, mt.GT.n_alt_alleles(), [1.0, mt.c1])
    assert np.isfinite(lin.beta).sum() > 0.9 * M


def test_plugin_config_entry_equals_the_public_call():
    """the config dict of statgen.py:394-401, applied through the registry (plugin.matrix_to_table_apply),
    gives the rows of the public call; Single returns array-valued statistics (LR:26-34), the wrapper unwraps them."""
    hb = _hb()
    from hail_b200 import plugin
    N, M = 500, 40
    rng = np.random.default_rng(21)
    x = rng.integers(0, 3, size=(M, N)).astype(np.float64)
    x[rng.random(x.shape) < 0.05] = np.nan
    mt = _mt_from_dosage(x, y=rng.normal(size=N), y2=rng.normal(size=N), c=rng.normal(size=N))
    ht = hb.linear_regression_rows(y=mt.y, x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.c], pass_through=["qual"])
    sel = mt.annotate_cols(__y_0=mt.y, __cov0=1.0, __cov1=mt.c).annotate_entries(__x=mt.GT.n_alt_alleles())
    cfg = {"name": "LinearRegressionRowsSingle", "yFields": ["__y_0"], "xField": "__x", "covFields": ["__cov0", "__cov1"],
           "rowBlockSize": 16, "passThrough": ["qual"]}
    raw = plugin.matrix_to_table_apply(sel, cfg)
    assert raw.row == plugin.lookup_matrix_to_table(cfg).typ(sel) == ht.row
    assert raw.beta.shape == (M, 1) and np.array_equal(raw.beta[:, 0], ht.beta, equal_nan=True)
    assert np.array_equal(raw.p_value[:, 0], ht.p_value, equal_nan=True) and np.array_equal(raw.n, ht.n)
    ch = plugin.matrix_to_table_apply(sel.annotate_cols(__y_1_0=mt.y2, __y_0_0=mt.y),
                                      {**cfg, "name": "LinearRegressionRowsChained", "yFields": [["__y_0_0"], ["__y_1_0"]]})
    assert ch.n.shape == (M, 2) and np.array_equal(ch.beta[0][:, 0], ht.beta, equal_nan=True)
