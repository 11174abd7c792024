"""GPU parity of `hwe_normalized_pca` (block Lanczos over lrr_run + lrr_at_times) against the exact-SVD oracle and the
reference's own numpy recipe (hail/python/test/hail/methods/test_pca.py:12-68, 101)."""
import numpy as np
import pytest

from oracle import pca_oracle as P

pytestmark = pytest.mark.gpu


def _hb():
    import hail_b200 as hb
    return hb


def _mt(x, **cols):
    hb = _hb()
    gt = hb.PackedGenotypes.from_dosage(np.where(np.isnan(x), -1, x).astype(np.int8))
    rows = {"locus": np.array([("1", i + 1) for i in range(x.shape[0])], dtype=object)}
    cols = dict(cols)
    cols["s"] = np.array([f"s{j}" for j in range(x.shape[1])], dtype=object)
    return hb.MatrixTable(gt, rows=rows, cols=cols, row_key=("locus",), col_key=("s",))


def _same_up_to_sign(a, b, rtol, atol=1e-10):
    a, b = np.asarray(a), np.asarray(b)
    for c in range(a.shape[1]):
        sgn = 1.0 if np.dot(a[:, c], b[:, c]) >= 0 else -1.0
        np.testing.assert_allclose(a[:, c], sgn * b[:, c], rtol=rtol, atol=atol)


def test_at_times_matches_numpy():
    import ctypes
    import torch
    hb = _hb()
    from hail_b200 import _lib
    rng = np.random.default_rng(1)
    M, N, L = 517, 1300, 7
    x = rng.integers(0, 3, size=(M, N)).astype(np.float64)
    x[rng.random((M, N)) < 0.05] = np.nan
    g = hb.PackedGenotypes.from_dosage(np.where(np.isnan(x), -1, x).astype(np.int8))
    coef = rng.normal(size=(M, 4))
    coef[:, 3] = 0.0
    t = rng.normal(size=(M, L))
    a = np.where(np.isnan(x), 0.0, np.take_along_axis(coef, np.nan_to_num(x).astype(np.int64), axis=1))
    want = a.T @ t
    dev = g.device
    ctx = _lib.context(dev.index)
    for n_splits in (1, 5):
        out = torch.empty((n_splits, N, L), dtype=torch.float64, device=dev)
        d_coef, d_t = torch.from_numpy(coef).to(dev), torch.from_numpy(t).to(dev)
        ctx.check(ctx.lib.lrr_at_times(ctx.handle, g.data.data_ptr(), M, g.stride, N, d_coef.data_ptr(), d_t.data_ptr(), L,
                                       n_splits, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        got = out.sum(dim=0).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-11)


def test_tiny_matrix_reference_recipe():   # test_pca.py:28-68
    hb = _hb()
    x = np.array([[1.0, np.nan, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 2.0, 0.0]])
    mt = _mt(x)
    ev, scores, loadings = hb.hwe_normalized_pca(mt.GT, k=3, compute_loadings=True)

    def normalize(a):
        ms = np.mean(a, axis=0, keepdims=True)
        return np.divide(np.subtract(a, ms), np.sqrt(2.0 * np.multiply(ms / 2.0, 1 - ms / 2.0) * a.shape[1]))

    g = np.pad(np.diag([1.0, 1, 2]), ((0, 1), (0, 0)), mode="constant")
    g[1, 0] = 1.0 / 3
    U, s, V = np.linalg.svd(normalize(g), full_matrices=0)
    np.testing.assert_allclose(ev, s * s, rtol=1e-5)
    np.testing.assert_allclose(np.abs(scores.scores), np.abs(U.dot(np.diag(s))), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(np.abs(loadings.loadings), np.abs(V.transpose()), rtol=1e-5, atol=1e-9)
    assert scores.count() == 4 and loadings.count() == 3 and list(scores.s) == ["s0", "s1", "s2", "s3"]


def test_bn_vs_exact_svd_oracle():
    hb = _hb()
    N, M, k = 700, 2500, 5
    bn = hb.balding_nichols_model(4, N, M, missing_rate=0.02, seed=7)
    x = bn.genotypes.to_dosage().astype(np.float64)
    x[x < 0] = np.nan
    x[11] = 0.0                      # monomorphic rows are dropped (pca.py:19)
    x[12] = np.where(np.isnan(x[12]), np.nan, 2.0)
    mt = _mt(x)
    ev, scores, loadings = hb.hwe_normalized_pca(mt.GT, k=k, compute_loadings=True)
    w_ev, w_scores, w_loadings, keep = P.hwe_normalized_pca(x, k)
    assert len(ev) == k and scores.count() == N and loadings.count() == int(keep.sum()) == M - 2
    np.testing.assert_allclose(ev, w_ev, rtol=1e-6)
    # 4 populations: 3 structure components stand clear of the bulk -> their vectors are well conditioned
    _same_up_to_sign(scores.scores[:, :3], w_scores[:, :3], rtol=1e-5, atol=1e-7)
    _same_up_to_sign(loadings.loadings[:, :3], w_loadings[:, :3], rtol=1e-5, atol=1e-7)
    # all k: projections (test_pca.py:101: A @ loadings == scores) and orthonormal loadings
    a, _ = P.hwe_normalize(x)
    # (the bulk components 4-5 are Ritz pairs: eigenvalue error ~ (vector error)^2, so vectors are looser than values)
    np.testing.assert_allclose(a.T @ loadings.loadings, scores.scores, rtol=1e-5, atol=5e-6)
    assert scores.n_iterations < 40
    np.testing.assert_allclose(loadings.loadings.T @ loadings.loadings, np.eye(k), atol=1e-8)
    assert [tuple(r) for r in loadings.locus] == [("1", i + 1) for i in range(M) if keep[i]]
    _, _, none = hb.hwe_normalized_pca(mt.GT, k=2)
    assert none is None


def test_filtered_columns_and_errors():
    hb = _hb()
    N, M = 400, 900
    bn = hb.balding_nichols_model(3, N, M, missing_rate=0.01, seed=3)
    x = bn.genotypes.to_dosage().astype(np.float64)
    x[x < 0] = np.nan
    mt = _mt(x)
    rng = np.random.default_rng(0)
    keep_cols = rng.random(N) < 0.7
    sub = mt.filter_cols(keep_cols)
    ev, scores, _ = hb.hwe_normalized_pca(sub.GT, k=2)
    w_ev, w_scores, _, _ = P.hwe_normalized_pca(x[:, keep_cols], 2)
    np.testing.assert_allclose(ev, w_ev, rtol=1e-6)
    _same_up_to_sign(scores.scores, w_scores, rtol=1e-5, atol=1e-7)
    assert list(scores.s) == [f"s{j}" for j in np.nonzero(keep_cols)[0]]
    with pytest.raises(hb.FatalError, match="requested invalid number of components"):
        hb.hwe_normalized_pca(mt.GT, k=0)
    mono = _mt(np.zeros((5, 50)))
    with pytest.raises(hb.FatalError, match="found 0 variants after filtering out monomorphic sites"):
        hb.hwe_normalized_pca(mono.GT, k=2)
    with pytest.raises(hb.ExpressionException):
        hb.hwe_normalized_pca(mt.GT.n_alt_alleles(), k=2)


def test_pca_at_400k_samples_vs_exact_eigensystem():
    """The BASELINE sample count: 400k samples x 1,536 variants from four populations.  The oracle's exact SVD of a
    1,536 x 400k matrix is replaced by the eigensystem of the small Gram matrix A A' (same eigenvalues; loadings = its
    eigenvectors, scores = A' U) built from the oracle's own hwe_normalize."""
    hb = _hb()
    N, M, k = 400_000, 1536, 3
    mt = hb.balding_nichols_model(4, N, M, missing_rate=0.01, seed=29)
    dos = mt.genotypes.to_dosage().astype(np.float32)
    dos[dos < 0] = np.nan
    a, keep = P.hwe_normalize(dos)                       # float64 [m, N]
    w, u = np.linalg.eigh(a @ a.T)
    order = np.argsort(w)[::-1][:k]
    want_ev, want_load = w[order], u[:, order]
    want_scores = a.T @ want_load
    ev, scores, loadings = hb.hwe_normalized_pca(mt.GT, k=k, compute_loadings=True)
    np.testing.assert_allclose(ev, want_ev, rtol=1e-6)
    assert want_ev[k - 1] > 3 * w[np.argsort(w)[::-1][k]]          # k structure components, well separated from the bulk
    _same_up_to_sign(scores.scores, want_scores, rtol=1e-5, atol=1e-5 * np.abs(want_scores).max())
    got_load = np.asarray(loadings.loadings)[keep] if np.asarray(loadings.loadings).shape[0] == M else np.asarray(loadings.loadings)
    _same_up_to_sign(got_load, want_load, rtol=1e-5, atol=1e-5 * np.abs(want_load).max())
