"""CPU restatement of `hl.hwe_normalized_pca` (exact SVD).  TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this.

Follows
  hail/python/hail/methods/pca.py:15-33      hwe_normalize: AC / n_called per variant, monomorphic variants dropped
                                             (0 < AC < 2 n_called), entry = (gt - mean) / sqrt(mean (2 - mean) m / 2),
                                             missing calls -> 0
  hail/hail/src/is/hail/methods/PCA.scala:34-118   SVD of the variants x samples matrix: eigenvalues = s^2,
                                             scores = V S per sample, loadings = U per (kept) variant
Pinned by the reference's own numpy check (hail/python/test/hail/methods/test_pca.py:28-68: tiny_m.vcf, a 3 x 4 matrix
written out in the test) -- tests/test_oracle_pca.py.
"""
import numpy as np


def hwe_normalize(x):
    """x [M, N] dosages, NaN = missing.  Returns (A [m, N] normalised entries of the kept variants, keep [M] bool)."""
    x = np.asarray(x, dtype=np.float64)
    called = ~np.isnan(x)
    ac = np.where(called, x, 0.0).sum(axis=1)
    n_called = called.sum(axis=1)
    keep = (ac > 0) & (ac < 2 * n_called)
    m = int(keep.sum())
    if m == 0:
        raise ValueError("hwe_normalize: found 0 variants after filtering out monomorphic sites.")
    mean = ac[keep] / n_called[keep]
    sd = np.sqrt(mean * (2.0 - mean) * m / 2.0)
    a = (x[keep] - mean[:, None]) / sd[:, None]
    return np.where(np.isnan(a), 0.0, a), keep


def pca(a, k):
    """a [m, N] (variants x samples).  Returns eigenvalues [k], scores [N, k], loadings [m, k]."""
    u, s, vt = np.linalg.svd(a, full_matrices=False)
    return (s * s)[:k], (vt.T * s)[:, :k], u[:, :k]


def hwe_normalized_pca(x, k=10):
    a, keep = hwe_normalize(x)
    ev, scores, loadings = pca(a, k)
    return ev, scores, loadings, keep
