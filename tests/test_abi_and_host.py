"""CPU-only tests: the C-ABI library loads and exports every declared symbol; host-side logic of the operator."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lrr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lrr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from hail_b200 import _lib, build

    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lrr_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in hail_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)
    lib.lrr_version.restype = ctypes.c_char_p
    assert lib.lrr_version().startswith(b"lrr_b200")
    lib.lrr_packed_stride.restype = ctypes.c_int64
    lib.lrr_packed_stride.argtypes = [ctypes.c_int64]
    assert lib.lrr_packed_stride(400000) == 100096 and lib.lrr_packed_stride(1) == 128 and lib.lrr_packed_stride(512) == 128


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hail_b200 import _lib

    with pytest.raises(_lib.LrrError, match="no CPU fallback"):
        _lib.Context(0)


class _FakeGenotypes:
    def __init__(self, m, n):
        self.n_variants, self.n_samples = m, n


def _fake_mt(m=10, n=8):
    import hail_b200 as hb

    rng = np.random.default_rng(0)
    rows = {"locus": np.arange(m), "alleles": np.arange(m), "filters": [set() for _ in range(m)],
            "qual": rng.normal(size=m), "foo": {"bar": rng.normal(size=m)}}
    cols = {"pheno": rng.normal(size=n), "c1": rng.normal(size=n)}
    return hb.MatrixTable(_FakeGenotypes(m, n), rows=rows, cols=cols, row_key=("locus", "alleles"))


def test_argument_validation_matches_reference():  # statgen.py:351-355, 195-224; TS:95-132
    import hail_b200 as hb

    mt = _fake_mt()
    x = mt.GT.n_alt_alleles()
    with pytest.raises(ValueError, match="found no values for 'y'"):
        hb.linear_regression_rows([], x, [1.0])
    with pytest.raises(ValueError, match="found empty inner list for 'y'"):
        hb.linear_regression_rows([[mt.pheno], []], x, [1.0])
    with pytest.raises(ValueError, match="no row field"):
        hb.linear_regression_rows(mt.pheno, x, [1.0], pass_through=["nope"])
    with pytest.raises(ValueError, match="duplicated field"):
        hb.linear_regression_rows(mt.pheno, x, [1.0], pass_through=["qual", "qual"])
    with pytest.raises(ValueError, match="not complex expressions"):
        hb.linear_regression_rows(mt.pheno, x, [1.0], pass_through=[mt.filters.length()])
    with pytest.raises(hb.ExpressionException):
        hb.linear_regression_rows(mt.pheno, x, [1.0], pass_through=[mt.pheno])
    # weights: statgen.py:437-467
    with pytest.raises(ValueError, match="When y is a list of lists, weights should be a list."):
        hb.linear_regression_rows([[mt.pheno]], x, [1.0], weights=mt.c1)
    with pytest.raises(ValueError, match="When y is a single list, weights should be a single expression."):
        hb.linear_regression_rows([mt.pheno], x, [1.0], weights=[mt.c1])
    with pytest.raises(ValueError, match="When y is a single expression, weights should be a single expression."):
        hb.linear_regression_rows(mt.pheno, x, [1.0], weights=[mt.c1])
    with pytest.raises(ValueError, match="Must specify same number of weights as groups of phenotypes"):
        hb.linear_regression_rows([[mt.pheno], [mt.pheno]], x, [1.0], weights=[mt.c1])
    with pytest.raises(hb.ExpressionException):
        hb.linear_regression_rows(mt.pheno, mt.pheno, [1.0])
    # key fields pass silently; nested fields keep their leaf name (TS:118-126)
    f = hb._get_regression_row_fields(mt, ["filters", mt.foo.bar, mt.qual, "locus", "alleles"], "linear_regression_rows")
    assert list(f) == ["filters", "bar", "qual"]


def test_intercept_warning():  # statgen.py:4881-4888
    import hail_b200 as hb

    mt = _fake_mt()
    with pytest.warns(UserWarning, match="no intercept"):
        assert hb._warn_if_no_intercept("linear_regression_rows", [mt.c1])
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert not hb._warn_if_no_intercept("linear_regression_rows", [1.0, mt.c1])


@pytest.mark.parametrize("intercept_pos", [None, 0, 2])
def test_group_basis_matches_oracle_prologue(intercept_pos):
    """The residualised basis handed to the device reproduces LR:65-78 (Q Q^T, Qty, yyp) whatever the rotation."""
    from hail_b200.statgen import FatalError, GroupBasis
    from oracle import linreg_oracle as O

    rng = np.random.default_rng(3)
    N, K, P = 200, 4, 3
    cov = rng.normal(size=(N, K)) + 1.0
    if intercept_pos is not None:
        cov[:, intercept_pos] = 1.0
    ys = rng.normal(size=(N, P)) + 5.0
    ys[rng.random((N, P)) < 0.05] = np.nan
    cov[rng.random((N, K)) < 0.02] = np.nan
    b = GroupBasis(ys, cov, np.arange(N))
    y, c, idx = O.complete_samples(ys, cov)
    n, k, d, Qt, Qty, yyp = O.prologue(y, c)
    assert (b.n, b.K, b.P, b.d) == (n, k, P, d) and np.array_equal(b.complete_idx, idx)
    assert b.has_intercept == (intercept_pos is not None)
    q_full = np.column_stack(([np.full(n, 1 / np.sqrt(n))] if b.has_intercept else []) + [b.q_cols.T])
    assert np.allclose(q_full.T @ q_full, np.eye(K), atol=1e-12)
    assert np.allclose(q_full @ q_full.T, Qt.T @ Qt, atol=1e-12)          # same projector
    assert np.allclose(b.yyp, yyp, rtol=1e-10)
    assert np.allclose(q_full @ b.qty, Qt.T @ Qty, atol=1e-10)            # same projection of y
    assert np.allclose(b.y_res.T, y - Qt.T @ Qty, atol=1e-10)
    assert np.abs(q_full.T @ b.y_res.T).max() < 1e-11
    if b.has_intercept:
        assert np.abs(b.q_cols.sum(axis=1)).max() < 1e-12
    with pytest.raises(FatalError, match="degrees of freedom"):
        GroupBasis(ys[:5], cov[:5], np.arange(5))
    with pytest.raises(FatalError, match="No complete samples"):
        GroupBasis(np.full((N, 1), np.nan), cov, np.arange(N))


@pytest.mark.parametrize("case", ["gauss+1", "no_intercept", "uncentred", "affine_span", "const_only", "collinear",
                                  "ill_conditioned", "planes_view"])
def test_transposed_prologue_equals_the_general_construction(case):
    """`orthonormal_basis_t` (the K x K-composed form the streamed call runs next to its host -> device copies) spans the same
    space as `orthonormal_basis`, finds the intercept in the same cases, and keeps row 0 exactly constant with the other
    rows summing to zero; rank-deficient and ill-conditioned covariates fall back to the general route."""
    from hail_b200.statgen import orthonormal_basis, orthonormal_basis_t
    rng = np.random.default_rng(11)
    n = 5000
    z = rng.normal(size=(n, 6))
    c = {"gauss+1": np.column_stack([np.ones(n), z]),
         "no_intercept": z,
         "uncentred": z + 5.0,
         "affine_span": np.column_stack([z[:, 0] + 3, z[:, 1] - 2 * z[:, 0] + 1, z[:, 2], np.full(n, 7.0)]),
         "const_only": np.full((n, 1), 3.0),
         "collinear": np.column_stack([np.ones(n), z[:, 0], 2 * z[:, 0] + 1]),
         "ill_conditioned": np.column_stack([np.ones(n), z[:, 0], z[:, 0] + 1e-6 * z[:, 1]]),
         "planes_view": np.stack([np.ones(n)] + [z[:, i] for i in range(6)]).T}[case]   # F-ordered view, as _execute passes it
    q, has = orthonormal_basis(c)
    q_t, has_t = orthonormal_basis_t(c)
    assert has_t == has and q_t.shape == (q.shape[1], n) and q_t.flags["C_CONTIGUOUS"]
    x = rng.normal(size=(n, 3))
    assert np.abs(q @ (q.T @ x) - q_t.T @ (q_t @ x)).max() < 1e-11                 # same projector
    assert np.abs(q_t @ q_t.T - np.eye(q_t.shape[0])).max() < 1e-12
    if has_t:
        assert np.all(q_t[0] == 1.0 / np.sqrt(n))
        if q_t.shape[0] > 1:
            assert np.abs(q_t[1:].sum(axis=1)).max() < 1e-12


def test_bn_parameters_are_seeded_and_shardable():
    from hail_b200 import bn

    pop, th, af = bn.bn_parameters(3, 500, 1000, missing_rate=0.25, seed=5)
    pop2, th2, af2 = bn.bn_parameters(3, 500, 1000, missing_rate=0.25, seed=5)
    assert np.array_equal(pop, pop2) and np.array_equal(th, th2)
    _, th3, _ = bn.bn_parameters(3, 500, 300, missing_rate=0.25, seed=5, first_variant=650)
    assert np.array_equal(th3, th[650:950])
    assert (th[:, :, 0] == round(0.25 * 65536)).all() and (np.diff(th.astype(np.int64), axis=2) >= 0).all()
    assert 0.0 < af.min() and af.max() < 1.0 and abs(af.mean() - 0.5) < 0.02  # Beta drift around U(0.1,0.9), F_st 0.1


def test_bn_mirror_distribution():
    from hail_b200 import bn
    from tests.bn_mirror import bn_fill_numpy

    pop, th, af = bn.bn_parameters(3, 4000, 50, seed=1)
    d = bn_fill_numpy(th, pop, 4000, seed=1)
    assert d.min() >= 0 and d.max() == 2
    freq = d.mean(axis=1) / 2.0
    assert np.abs(freq - af[:, pop].mean(axis=1)).max() < 0.03


def test_neighbour_calls_validate_before_any_device_work():
    """logistic_regression_rows / hwe_normalized_pca raise the reference's errors on the host (no GPU needed):
    statgen.py:960-984, LogisticRegression.scala:44-61, PCA.scala:35-37."""
    import hail_b200 as hb

    mt = _fake_mt()
    x = mt.GT.n_alt_alleles()
    yb = (mt.pheno.values > 0).astype(np.float64)
    mt2 = mt.annotate_cols(yb=yb, q=mt.pheno.values, one=np.ones(8))
    x2 = mt2.GT.n_alt_alleles()
    with pytest.raises(TypeError):
        hb.logistic_regression_rows("rao", mt2.yb, x2, [1.0])
    with pytest.raises(ValueError, match="at least one covariate"):
        hb.logistic_regression_rows("wald", mt2.yb, x2, [])
    with pytest.raises(ValueError, match="found no values for 'y'"):
        hb.logistic_regression_rows("lrt", [], x2, [1.0])
    with pytest.raises(hb.ExpressionException):
        hb.logistic_regression_rows("firth", mt2.yb, mt2.yb, [1.0])
    with pytest.raises(hb.FatalError, match="equal to 0 or 1"):
        hb.logistic_regression_rows("wald", mt2.q, x2, [1.0])
    with pytest.raises(hb.FatalError, match="must be non-constant"):
        hb.logistic_regression_rows("wald", mt2.one, x2, [1.0])
    with pytest.raises(hb.FatalError, match="degrees of freedom"):
        hb.logistic_regression_rows("wald", mt2.yb, x2, [1.0] + [mt2.c1] * 7)
    with pytest.raises(hb.FatalError, match="requested invalid number of components"):
        hb.hwe_normalized_pca(mt.GT, k=0)
    with pytest.raises(hb.ExpressionException):
        hb.hwe_normalized_pca(x, k=2)
    with pytest.raises(TypeError):
        hb.hwe_normalized_pca(mt.GT, k=2.0)


# ---------------------------------------------------------------------------------------------
# the plugin entry: config dict -> relational function, by name (RelationalFunctions.scala:112-138)
def test_plugin_lookup_by_config_name_and_schema():
    import json

    import pytest

    from hail_b200 import plugin

    cfg = {"name": "LinearRegressionRowsSingle", "yFields": ["__y_0", "__y_1"], "xField": "__uid_x", "covFields": ["__cov0"],
           "rowBlockSize": 16, "passThrough": ["rsid"]}
    f = plugin.lookup_matrix_to_table(cfg)
    assert isinstance(f, plugin.LinearRegressionRowsSingle) and f.preserves_partition_counts()
    assert isinstance(plugin.lookup_matrix_to_table(json.dumps(cfg)), plugin.LinearRegressionRowsSingle)   # the JSON the JVM sees

    class _G:   # a MatrixTable needs only its shape here
        n_variants, n_samples = 3, 4
    mt = plugin.MatrixTable(_G(), rows={"locus": [1, 2, 3], "rsid": ["a", "b", "c"]}, row_key=("locus",))
    # result schema: key, pass-through, then the statistics in the reference's order (LR:26-42)
    assert f.typ(mt) == ["locus", "rsid", "n", "sum_x", "y_transpose_x", "beta", "standard_error", "t_stat", "p_value"]
    ch = plugin.lookup_matrix_to_table({**cfg, "name": "LinearRegressionRowsChained", "yFields": [["__y_0_0"], ["__y_1_0", "__y_1_1"]]})
    assert isinstance(ch, plugin.LinearRegressionRowsChained) and ch.chained
    with pytest.raises(ValueError, match="no MatrixToTableFunction registered"):
        plugin.lookup_matrix_to_table({**cfg, "name": "PoissonRegression"})
    with pytest.raises(ValueError, match="bad config"):
        plugin.lookup_matrix_to_table({k: v for k, v in cfg.items() if k != "covFields"})
    with pytest.raises(ValueError, match="yFields"):
        plugin.lookup_matrix_to_table({**cfg, "name": "LinearRegressionRowsChained"})   # chained needs lists of lists
    with pytest.raises(KeyError, match="no column field"):
        f.execute(mt)
