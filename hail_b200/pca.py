"""`hwe_normalized_pca` on the B200 path -- the step that makes the PC covariates of a GWAS.

Mirrors
  * hail/python/hail/methods/pca.py:15-33     hwe_normalize: per variant AC and n_called over the kept samples,
                                              monomorphic variants dropped (0 < AC < 2 n_called; FatalError when none is
                                              left), entry = (gt - mean) / sqrt(mean (2 - mean) m / 2), missing -> 0
  * hail/python/hail/methods/pca.py:35-98     hwe_normalized_pca(call_expr, k, compute_loadings) ->
                                              (eigenvalues, scores table keyed by the column key, loadings table keyed
                                              by the row key or None)
  * hail/hail/src/is/hail/methods/PCA.scala:34-118   k < 1 and "Found only N non-zero eigenvalues" are fatal;
                                              eigenvalues = s^2, scores = V S, loadings = U of the variants x samples matrix
The reference hands the matrix to Spark's ARPACK SVD (or, off Spark, to a block Krylov iteration, pca.py:345-424).  Here
the normalised matrix A is never materialised: a block Lanczos iteration with full re-orthogonalisation needs only
    T = A V      the per-variant sweep over the packed genotypes (lrr_run with the columns of V as phenotypes:
                 y_transpose_x = sum_j x_imputed[v, j] V[j, c], exact-integer tensor-core kernels), then
                 T[v] = (ytx[v] - mean_v colsum(V)) / sd_v
    W = A' T     lrr_at_times (csrc/gram_kernel.cu) with the per-variant table of normalised entry values
per step, i.e. two passes over the 2-bit genotypes; the small dense algebra (QR of n x L blocks, the Rayleigh-Ritz
eigenproblem) runs in torch on the device.  Iterations stop when the top-k Ritz values move by less than `_tol`
(relative) or the Krylov space is exhausted.
"""
from __future__ import annotations

import logging
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .matrixtable import CallExpression, ExpressionException, Table
from .statgen import FatalError, _run_device

log = logging.getLogger("hail_b200")


class _DeviceColumns:
    """Quacks like the numpy arrays `statgen._add_group` hands to lrr_add_group (`.ctypes.data`, `.size`), but the
    pointer is a device pointer: the library copies with cudaMemcpyDefault, so the columns never visit the host."""

    def __init__(self, t):
        self.tensor = t.contiguous()
        self.size = self.tensor.numel()
        self.ctypes = SimpleNamespace(data=self.tensor.data_ptr())


def _sweep_basis(idx32, cols):
    """A `linear_regression_rows` group with no covariates whose phenotypes are the given columns [P, n]
    (a numpy array or a float64 CUDA tensor)."""
    P, n = cols.shape
    y = _DeviceColumns(cols) if isinstance(cols, torch.Tensor) else np.ascontiguousarray(cols, dtype=np.float64)
    return SimpleNamespace(n=n, K=0, P=P, has_intercept=False, complete_idx=idx32, q_cols=np.empty(0), qty=np.empty(0),
                           y_res=y, yyp=np.ones(P), weighted=False)


def _qr(W):
    """Thin QR of a tall block.  Cholesky-QR applied twice (two small Gram matrices and triangular solves: the
    Householder QR of a 400k x 16 block costs more than a pass over the genotypes); falls back to Householder QR when
    the block is numerically rank deficient."""
    try:
        R1 = torch.linalg.cholesky(W.t() @ W, upper=True)
        Q = torch.linalg.solve_triangular(R1, W, upper=True, left=False)
        R2 = torch.linalg.cholesky(Q.t() @ Q, upper=True)
        Q = torch.linalg.solve_triangular(R2, Q, upper=True, left=False)
        R = R2 @ R1
        if bool(torch.isfinite(R).all()) and float(R.diagonal().abs().min()) > 1e-8 * float(R.diagonal().abs().max()):
            return Q, R
    except Exception:   # not positive definite: rank deficient block
        pass
    return torch.linalg.qr(W)


def _allreduce(t, sharded):
    """Sum over the ranks of a variant-sharded run (NCCL all-reduce over NVLink); identity otherwise."""
    if sharded:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def hwe_normalized_pca(call_expr, k=10, compute_loadings=False, *, _oversample=8, _tol=1e-9, _max_iterations=60,
                       _restart_blocks=6, _seed=0, _sharded=False):
    """Run principal component analysis (PCA) on the Hardy-Weinberg-normalized genotype call matrix
    (drop-in for `hl.hwe_normalized_pca`, pca.py:35).  Returns (eigenvalues, scores, loadings).

    `_sharded=True` (one process per GPU, torch.distributed initialised): every rank passes its own contiguous range of
    VARIANTS over the same samples.  Unlike the regression sweep this algorithm has a real exchange step: the
    transposed product A' T and the small Gram matrices T' T are sums over variants, so they are all-reduced (n x L
    and L x L float64 per iteration); eigenvalues and scores come out identical on every rank, loadings stay sharded."""
    if not isinstance(call_expr, CallExpression):
        raise ExpressionException("'hwe_normalized_pca/call_expr': expected a call expression (e.g. mt.GT)")
    if not isinstance(k, int) or isinstance(k, bool):
        raise TypeError("hwe_normalized_pca: 'k' must be int")
    mt = call_expr.source
    if k < 1:   # PCA.scala:35-37
        raise FatalError(f"requested invalid number of components: {k}\n  Expect componenents >= 1")
    from .genotypes import HostBedGenotypes
    g = mt.genotypes
    if isinstance(g, HostBedGenotypes):
        g = g.to_device()
    dev = g.device
    ctx = _lib.context(dev.index)
    M, N = g.n_variants, g.n_samples
    idx = np.ascontiguousarray(np.asarray(mt.col_index), dtype=np.int32)
    n = idx.size
    f64 = dict(dtype=torch.float64, device=dev)

    with torch.cuda.device(dev):
        # ---- hwe_normalize (pca.py:15-33): counts from one sweep against a constant column ----
        o = _run_device(g, [_sweep_basis(idx, np.ones((1, n)))], guard=False)[0]
        n_called = (n - o["n_missing"]).to(torch.float64)
        mean = o["sum_x"] / float(n)                  # the mean-imputed column sums to n * mean
        keep = (mean > 0.0) & (mean < 2.0) & (n_called > 0)
        m = int(_allreduce(keep.sum().to(torch.int64).reshape(1), _sharded).item())
        if m == 0:
            raise FatalError("hwe_normalize: found 0 variants after filtering out monomorphic sites.")
        mean = torch.where(keep, mean, torch.zeros_like(mean))
        inv_sd = torch.where(keep, torch.rsqrt(torch.clamp(mean * (2.0 - mean) * (m / 2.0), min=1e-300)), torch.zeros_like(mean))
        codes = torch.tensor([0.0, 1.0, 2.0], **f64)
        coef = torch.zeros((M, 4), **f64)
        coef[:, :3] = (codes[None, :] - mean[:, None]) * inv_sd[:, None]       # missing call -> 0 (pca.py:30)
        d_idx = torch.from_numpy(idx.astype(np.int64)).to(dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        strips = (g.stride + 127) // 128
        n_splits = int(max(1, min(64, -(-8 * 148 // strips), M // 32 or 1)))

        def a_times(V):            # [n, L] -> [M, L]
            ytx = _run_device(g, [_sweep_basis(idx, V.t().contiguous())], guard=False)[0]["y_transpose_x"]
            T = (ytx - mean[:, None] * V.sum(dim=0)[None, :]) * inv_sd[:, None]
            return torch.where(keep[:, None], T, torch.zeros_like(T))

        def at_times(T):           # [M, L] -> [n, L]
            L = T.shape[1]
            out = torch.empty((n_splits, N, L), **f64)
            Tc = T.contiguous()
            ctx.check(ctx.lib.lrr_at_times(ctx.handle, g.data.data_ptr(), M, g.stride, N, coef.data_ptr(), Tc.data_ptr(), L,
                                           n_splits, out.data_ptr(), stream))
            return _allreduce(out.sum(dim=0)[d_idx].contiguous(), _sharded)

        def gram(T):               # T' T summed over every rank's variants
            return _allreduce(T.t() @ T, _sharded)

        # ---- block Lanczos with full re-orthogonalisation, Rayleigh-Ritz over the accumulated Krylov space ----
        L = int(min(n, -(-max(k + _oversample, k) // 4) * 4, 24))   # block width: a multiple of 4 (the kernels' column tiles)
        if min(n, m) < k:   # PCA.scala:47-51
            raise FatalError(f"Found only {min(n, m)} non-zero (or nearly zero) eigenvalues, but user requested {k} "
                             "principal components.")
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(_seed))
        V0 = _qr(torch.randn((n, L), generator=gen, **f64))[0].contiguous()
        if _sharded:
            import torch.distributed as dist
            dist.broadcast(V0, 0)
        Vs = [V0]
        Ts = []
        prev = None
        n_cols = L
        n_restarts = n_sweeps = 0
        converged = False
        for it in range(int(_max_iterations)):
            Ts.append(a_times(Vs[-1]))
            n_sweeps += 1
            Tall = torch.cat(Ts, dim=1)
            ritz = torch.linalg.eigvalsh(gram(Tall)).flip(0)[:k]
            if prev is not None and bool(((ritz - prev).abs() <= _tol * ritz.abs().clamp(min=1e-300)).all()):
                converged = True
                break
            prev = ritz
            if n_cols >= n and n_restarts == 0:        # the subspace is the whole sample space: Rayleigh-Ritz is exact
                converged = True
                break
            W = at_times(Ts[-1])
            scale = float(W.norm()) / max(W.shape[1], 1) ** 0.5
            Vall = torch.cat(Vs, dim=1)
            for _ in range(2):
                W = W - Vall @ (Vall.t() @ W)
            Q, R = _qr(W)
            good = R.diagonal().abs() > 1e-10 * max(scale, 1e-300)   # directions that are new (not roundoff of old ones)
            n_new = int(min(int(good.sum()), n - n_cols))
            if n_new == 0:                             # invariant subspace: the Krylov space is exhausted
                converged = True
                break
            Vs.append(Q[:, good][:, :n_new].contiguous())
            n_cols += n_new
            if len(Vs) > _restart_blocks and n_cols < n:
                # thick restart: keep the leading Ritz vectors of everything but the newest block (A Vr = Tall Wm needs
                # no sweep) so that the Rayleigh-Ritz problem and the re-orthogonalisation stay small
                Vold, Told = torch.cat(Vs[:-1], dim=1), torch.cat(Ts, dim=1)
                _, Wr = torch.linalg.eigh(gram(Told))
                Wr = Wr.flip(1)[:, :L]
                Vs = [Vold @ Wr, Vs[-1]]
                Ts = [Told @ Wr]
                n_cols = Vs[0].shape[1] + Vs[1].shape[1]
                n_restarts += 1
        if not converged:
            # components inside the noise bulk (eigenvalues a fraction of a percent apart) converge slowly in any Lanczos
            # method (Spark's ARPACK allows 300 iterations, mllib RowMatrix.computeSVD); the Ritz pairs are returned
            log.warning("hwe_normalized_pca: top-%d Ritz values still moving by more than %g after %d sweeps", k, _tol, n_sweeps)
        Vall = torch.cat(Vs[:len(Ts)], dim=1)
        Tall = torch.cat(Ts, dim=1)
        evals, Wm = torch.linalg.eigh(gram(Tall))
        evals, Wm = evals.flip(0)[:k], Wm.flip(1)[:, :k]
        if evals.numel() < k or bool((evals[:k] <= 1e-12 * evals[0].clamp(min=1e-300)).any()):   # PCA.scala:47-51
            nz = int((evals > 1e-12 * evals[0]).sum())
            raise FatalError(f"Found only {nz} non-zero (or nearly zero) eigenvalues, but user requested {k} principal components.")
        s = evals.sqrt()
        scores = (Vall @ Wm) * s[None, :]
        loadings = ((Tall @ Wm) / s[None, :])[keep] if compute_loadings else None
        ctx.check(ctx.lib.lrr_clear_groups(ctx.handle))
        eigenvalues = evals.cpu().numpy().tolist()
        scores_h = scores.cpu().numpy()
        loadings_h = loadings.cpu().numpy() if compute_loadings else None
        keep_h = keep.cpu().numpy()

    sf = OrderedDict()
    for kf in mt.col_key:
        sf[kf] = mt.col[kf]
    sf["scores"] = scores_h
    scores_t = Table(sf, key=mt.col_key, n_rows=n)
    scores_t.n_iterations = n_sweeps
    scores_t.converged = converged
    loadings_t = None
    if compute_loadings:
        lf = OrderedDict()
        for kf in mt.row_key:
            v = mt.row[kf]
            lf[kf] = v[keep_h] if isinstance(v, np.ndarray) else [x for x, kp in zip(v, keep_h) if kp]
        lf["loadings"] = loadings_h
        loadings_t = Table(lf, key=mt.row_key, n_rows=int(keep_h.sum()))   # this rank's kept variants when sharded
        loadings_t.kept_variants = keep_h
    return eigenvalues, scores_t, loadings_t
