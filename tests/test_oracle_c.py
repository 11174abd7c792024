"""The C restatement (oracle/linreg_oracle.c) against the numpy oracle and the golden fixtures."""
import os

import numpy as np
import pytest

from oracle import bed as obed
from oracle import c_oracle
from oracle import linreg_oracle as O
from tests.helpers import GOLDEN, assert_fields_close, load_regression_linear


def test_c_two_sided_p_matches_numpy():
    L = c_oracle.lib()
    for d in (1, 2, 7, 30, 994, 399989, 499989):
        for t in (0.0, 1e-9, 1e-3, 0.3, 1.0, 1.7, 1.75, 2.0, 3.3, 8.0, 20.0, 37.0):
            want = float(O.two_sided_p(t, d))
            got = L.lrr_oracle_two_sided_p(t, float(d))
            # CF near x~1 amplifies the rounding of x by ~a/1.5: 1e-11 relative at d=4e5 (tolerance on p is 1e-5)
            assert abs(got - want) <= 1e-9 * want + 1e-300, (d, t, got, want)
    assert np.isnan(L.lrr_oracle_two_sided_p(float("nan"), 5.0))


def test_c_oracle_on_reference_golden():
    x, y, cov, doc = load_regression_linear()
    covs = np.column_stack([np.ones(8), cov])
    rows = obed.encode_rows(x)
    assert np.array_equal(obed.decode_rows(rows, 8), x, equal_nan=True)
    got = c_oracle.linreg_group_bed(rows, 8, y[:, None], covs)
    exp = doc["expected"]["with_cov"]
    for pos in ("1", "2", "3"):
        for f, v in exp[pos].items():
            assert abs(got[f][int(pos) - 1, 0] - v) < 5e-7
    for v in exp["nan_se"]:
        assert np.isnan(got["standard_error"][v - 1, 0])


@pytest.mark.parametrize("P,K", [(1, 1), (2, 3), (3, 0)])
def test_c_oracle_matches_numpy_on_fastlmm(P, K):
    z = np.load(os.path.join(GOLDEN, "fastlmm.npz"))
    N, M = int(z["n_samples"]), int(z["n_variants"])
    rows = obed.bed_body(z["bed"], N, M)[:300]
    rng = np.random.default_rng(1)
    ys = np.column_stack([z["pheno"]] + [rng.normal(size=N) for _ in range(P - 1)])
    ys[rng.random(ys.shape) < 0.03] = np.nan
    cov = np.column_stack([np.ones(N), z["cov"], rng.normal(size=(N, 2))])[:, :K]
    x = obed.decode_rows(rows, N)
    want = O.linreg_group(x, ys, cov)
    got = c_oracle.linreg_group_bed(rows, N, ys, cov, n_threads=2)
    assert_fields_close(got, want, rel=1e-9, rel_p=1e-9, t_floor=1e-11, ctx=f"P={P} K={K}")


def test_bn_fixture_has_missing_and_decodes():
    z = np.load(os.path.join(GOLDEN, "bn_4x1024.npz"))
    rows = obed.bed_body(z["bed"], int(z["n_samples"]), int(z["n_variants"]))
    x = obed.decode_rows(rows, 4)
    assert x.shape == (1024, 4) and np.isnan(x).sum() > 0
    assert np.array_equal(obed.encode_rows(x), rows)
