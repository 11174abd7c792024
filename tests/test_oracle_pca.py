"""The PCA oracle against the reference's own numpy check (test_pca.py:28-68)."""
import numpy as np

from oracle import pca_oracle as P


def test_tiny_matrix_matches_reference_numpy_recipe():
    # tiny_m.vcf as the reference test writes it out: g = pad(diag([1, 1, 2])), g[1, 0] = 1/3 is the mean-imputed
    # missing call of variant 0 in sample 1 (samples x variants); i.e. variants x samples calls below
    x = np.array([[1.0, np.nan, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 2.0, 0.0]])
    ev, scores, loadings, keep = P.hwe_normalized_pca(x, k=3)
    assert keep.all()

    def normalize(a):   # verbatim recipe of test_pca.py:51-53
        ms = np.mean(a, axis=0, keepdims=True)
        return np.divide(np.subtract(a, ms), np.sqrt(2.0 * np.multiply(ms / 2.0, 1 - ms / 2.0) * a.shape[1]))

    g = np.pad(np.diag([1.0, 1, 2]), ((0, 1), (0, 0)), mode="constant")
    g[1, 0] = 1.0 / 3
    n = normalize(g)
    U, s, V = np.linalg.svd(n, full_matrices=0)
    np.testing.assert_allclose(ev, s * s, rtol=1e-5)
    np.testing.assert_allclose(np.abs(scores), np.abs(U.dot(np.diag(s))), rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(np.abs(loadings), np.abs(V.transpose()), rtol=1e-5, atol=1e-12)


def test_monomorphic_variants_are_dropped_and_all_monomorphic_is_fatal():
    x = np.array([[0.0, 0.0, 0.0], [2.0, 2.0, np.nan], [0.0, 1.0, 2.0], [1.0, np.nan, 0.0]])
    a, keep = P.hwe_normalize(x)
    assert keep.tolist() == [False, False, True, True] and a.shape == (2, 3) and a[1, 1] == 0.0
    import pytest
    with pytest.raises(ValueError, match="found 0 variants"):
        P.hwe_normalize(x[:2])
