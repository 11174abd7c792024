"""`linear_regression_rows` -- host side of the B200 path.

Mirrors the reference's operator interface for this one path:
  * Python API + validation     hail/python/hail/methods/statgen.py:195-224 (pass_through), 227-408
  * plugin config + schema      hail/python/hail/ir/table_ir.py:948-997; LinearRegression.scala:18-44, 198-224
  * driver prologue             LinearRegression.scala:47-78 / 228-257, RegressionUtils.scala:88-128
The per-partition hot loop (LR:95-193 / 274-402) runs in the CUDA library behind include/lrr_b200.h.
PyTorch is used only for device buffers / streams; there is no CPU fallback for the hot loop.
"""
from __future__ import annotations

import ctypes
import itertools
import time
import logging
import warnings
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .matrixtable import (ChainedField, ColumnExpression, EntryExpression, Expression, ExpressionException,
                          MatrixTable, RowExpression, Table)

log = logging.getLogger("hail_b200")

try:  # skinny (n x K) @ (K x K) products are 10x slower under multi-threaded OpenBLAS than on one thread
    from threadpoolctl import threadpool_limits as _blas_limits
except Exception:  # pragma: no cover
    import contextlib

    def _blas_limits(limits=None):
        return contextlib.nullcontext()

STAT_FIELDS = ["y_transpose_x", "beta", "standard_error", "t_stat", "p_value"]


class FatalError(Exception):
    """hail.utils.java.FatalError -- what `fatal(...)` in the JVM surfaces as (py4j_backend.py:302-307)."""


def _plural(n, word):
    return word if n == 1 else word + "s"


# ---- statgen.py:195-224 ------------------------------------------------------------------------
def _get_regression_row_fields(mt: MatrixTable, pass_through, method) -> "OrderedDict[str, object]":
    row_fields = OrderedDict((k, k) for k in mt.row_key)
    for f in pass_through:
        if isinstance(f, str):
            if f not in mt.row:
                raise ValueError(f"'{method}/pass_through': MatrixTable has no row field {f!r}")
            if f in row_fields:
                if f in mt.row_key:  # allow silent pass through of key fields
                    pass
                else:
                    raise ValueError(f"'{method}/pass_through': found duplicated field {f!r}")
            row_fields[f] = mt.row[f]
        else:
            assert isinstance(f, Expression)
            if not f.is_nested_field:
                raise ValueError(f"'{method}/pass_through': expect fields or nested fields, not complex expressions")
            if f.axes != frozenset({"row"}):
                raise ExpressionException(
                    f"'{method}/pass_through': require row-indexed fields, found indices {sorted(f.axes)}")
            name = f.name
            if name in row_fields:
                if not (name in mt.row_key and f.values is mt.row.get(name)):
                    raise ValueError(f"'{method}/pass_through': found duplicated field {name!r}")
            row_fields[name] = f.values
    for k in mt.row_key:
        del row_fields[k]
    return row_fields


# ---- statgen.py:4881-4888 ----------------------------------------------------------------------
def _warn_if_no_intercept(caller, covariates):
    if all(isinstance(e, Expression) and e.axes for e in covariates):
        warnings.warn(f"{caller}: model appears to have no intercept covariate."
                      "\n    To include an intercept, add 1.0 to the list of covariates.")
        return True
    return False


def _column_values(e, mt, what):
    """Evaluate a column-indexed float expression to float64 [n_cols] (NaN = missing)."""
    if isinstance(e, ColumnExpression):
        if e.source is not mt:
            raise ExpressionException(f"'{what}': expression is not from the same MatrixTable as 'x'")
        return e.values
    if isinstance(e, Expression):
        raise ExpressionException(f"'{what}': expected a column-indexed expression, found indices {sorted(e.axes)}")
    if isinstance(e, (bool, int, float, np.integer, np.floating)):
        return np.full(mt.count_cols(), float(e))
    raise TypeError(f"'{what}': expected expression of type float64, found {type(e).__name__}")


# ---- the driver prologue: RU:88-128 + LR:47-78 ---------------------------------------------------
class GroupBasis:
    """What the reference broadcasts per group (ChainedLinregInput, LR:409-417), in the residualised form the
    device consumes (include/lrr_b200.h lrr_add_group)."""

    def __init__(self, ys, cov, col_index, group_index=None, weights=None):
        ys = np.asarray(ys, dtype=np.float64)   # [n_cols, P]
        cov = np.asarray(cov, dtype=np.float64).reshape(ys.shape[0], -1)  # [n_cols, K]
        n_cols, P = ys.shape
        K = cov.shape[1]
        if P == 0:
            raise FatalError("No phenotypes present.")  # RU:97-98
        keep = ~np.isnan(ys).any(axis=1) & ~np.isnan(cov).any(axis=1)  # RU:100-110
        if weights is not None:   # a sample also needs its weight (statgen.py:539-543)
            weights = np.asarray(weights, dtype=np.float64)
            keep &= ~np.isnan(weights)
        n = int(keep.sum())
        if n == 0:
            raise FatalError("No complete samples: each sample is missing its phenotype or some covariate")  # RU:113-114
        if n < n_cols:
            log.warning("%d of %d samples have a missing phenotype or covariate.", n_cols - n, n_cols)  # RU:124-125
        d = n - K - 1
        if d < 1:  # LR:55-58
            raise FatalError(f"{n} samples and {K + 1} {_plural(K, 'covariate')} (including x) implies {d} degrees of freedom.")
        tag = "" if group_index is None else f"[{group_index}]"
        log.info("linear_regression_rows%s: running on %d samples for %d response %s y,\n"
                 "    with input variable x, and %d additional %s...", tag, n, P, _plural(P, "variable"), K,
                 _plural(K, "covariate"))  # LR:60-63 / 241-244
        self.weighted = weights is not None
        col_index = np.asarray(col_index)
        if n < n_cols:   # (every sample complete: no 32 MB copies of what is already there)
            ys, cov, col_index = ys[keep], cov[keep], col_index[keep]
            if self.weighted:
                weights = weights[keep]
        with _blas_limits(limits=1):
            if self.weighted:
                self._build_weighted(ys, cov, np.sqrt(weights), col_index, n, K, P, d)
            else:
                self._build(ys, cov, col_index, n, K, P, d)

    def _build_weighted(self, y, c, sw, kept_index, n, K, P, d):
        """statgen.py:557-581: y and the covariates are scaled by sqrt(w) before the QR; x is scaled on the device
        (the shipped columns carry a second factor sqrt(w), see lrr_add_group_weighted)."""
        self.n, self.K, self.P, self.d = n, K, P, d
        self.complete_idx = np.ascontiguousarray(kept_index, dtype=np.int32)
        y = y * sw[:, None]
        c = c * sw[:, None]
        if K > 0:
            q, ok = _gram_orthonormalise(c)
            if not ok or q.shape[1] != K:
                q, _ = np.linalg.qr(c, mode="reduced")
        else:
            q = np.zeros((n, 0))
        qty = q.T @ y
        self.has_intercept = False
        self.qty = np.ascontiguousarray(qty)
        self.yyp = np.ascontiguousarray(np.einsum("ij,ij->j", y, y) - np.einsum("ij,ij->j", qty, qty))
        y_res = y - q @ qty
        y_res -= q @ (q.T @ y_res)
        self.q_cols = np.ascontiguousarray((q * sw[:, None]).T)        # [K, n]
        self.y_res = np.ascontiguousarray((y_res * sw[:, None]).T)     # [P, n]
        self.sqrt_w = np.ascontiguousarray(sw)

    def _build(self, y, c, kept_index, n, K, P, d):
        self.n, self.K, self.P, self.d = n, K, P, d
        self.complete_idx = np.ascontiguousarray(kept_index, dtype=np.int32)  # into the packed store
        # everything below works on TRANSPOSED planes ([K, n], [P, n]: what the device consumes), so that no n x K array is
        # ever transposed or re-stacked: the prologue runs next to a host -> device copy that takes the memory bus
        if K > 0:
            q_t, has_intercept = orthonormal_basis_t(c)    # LR:67 (only Q Q^T enters the results)
        else:
            q_t, has_intercept = np.zeros((0, n)), False
        qty = q_t @ y                                  # LR:71  [K, P]
        self.has_intercept = has_intercept
        self.qty = np.ascontiguousarray(qty)
        self.yyp = np.ascontiguousarray(np.einsum("ij,ij->j", y, y) - np.einsum("ij,ij->j", qty, qty))  # LR:78
        y_res_t = y.T - qty.T @ q_t                    # so that y_res . x == ytx - Qty^T qtx (LR:146)  [P, n]
        y_res_t -= (q_t @ y_res_t.T).T @ q_t           # one re-orthogonalisation pass
        kd0 = 1 if has_intercept else 0
        self.q_cols = np.ascontiguousarray(q_t[kd0:])       # [Kd, n] (rows of a C-ordered array: no copy)
        self.y_res = np.ascontiguousarray(y_res_t)          # [P, n]


def _gram_orthonormalise(a, rank_tol=1e-11):
    """Orthonormal basis of span(a) from the K x K Gram matrix (two passes = CholeskyQR2-style accuracy).

    Returns (q [n, r], ok): ok is False when the columns are too ill-conditioned for the Gram route.
    """
    g = a.T @ a
    w, v = np.linalg.eigh(g)
    wmax = w.max() if w.size else 0.0
    if not np.isfinite(wmax) or wmax <= 0.0:
        return a[:, :0], True
    keep = w > rank_tol * wmax
    if w[keep].min() < 1e-7 * wmax:   # cond(a)^2 too large for one Gram pass to be safe
        return None, False
    q = a @ (v[:, keep] / np.sqrt(w[keep]))
    g2 = q.T @ q                      # second pass removes the O(cond^2 eps) loss of orthogonality
    w2, v2 = np.linalg.eigh(g2)
    q = q @ ((v2 / np.sqrt(w2)) @ v2.T)
    return q, True


def orthonormal_basis(c):
    """Orthonormal basis Q of the covariate column space, as `qr.reduced.justQ(cov)` provides in LR:67.

    Only the projector Q Q^T enters LR:139-146, so any orthonormal basis of the same space gives the same
    results.  When the constant vector lies in the span, column 0 is made exactly 1/sqrt(n) and the other K-1
    columns are orthogonal to it: the device then forms that column's projection exactly from integer genotype
    counts, which removes the dominant cancellation in x.x - |Q^T x|^2.  Returns (Q [n, K], has_intercept).
    """
    n, K = c.shape
    one = np.full(n, 1.0 / np.sqrt(n))
    # is the constant in the span?  (least squares on the small Gram system; verified by the residual)
    q_all, ok = _gram_orthonormalise(c)
    if not ok or q_all.shape[1] != K:
        q_all, _ = np.linalg.qr(c, mode="reduced")        # ill-conditioned or rank-deficient: Householder QR
    u = q_all.T @ one
    resid = one - q_all @ u
    if np.linalg.norm(resid) > 1e-9 or q_all.shape[1] != K:
        return q_all, False
    # rotate inside the span: centre the columns, orthonormalise what is left (rank K-1)
    cc = q_all - np.outer(one, u)                            # = (I - 1 1^T/n) q_all
    rest, ok = _gram_orthonormalise(cc, rank_tol=1e-9)
    if not ok or rest.shape[1] != K - 1:
        q2, _ = np.linalg.qr(cc, mode="reduced")
        w = np.abs(np.linalg.svd(cc, compute_uv=False))
        rest = q2[:, : K - 1] if K > 1 else q2[:, :0]
        if K > 1 and w[K - 2] < 1e-9:
            return q_all, False
    rest = rest - np.outer(one, one @ rest)                  # exact-zero column sums up to roundoff
    return np.column_stack([one, rest]), True


def orthonormal_basis_t(c):
    """`orthonormal_basis` in transposed form: returns (Q^T [K, n] C-contiguous, has_intercept).

    Same construction -- two Gram (CholeskyQR2-style) passes, then a rotation inside the span that makes row 0 exactly
    1/sqrt(n) and the other rows orthogonal to it -- but every step after the first pass is composed in K x K algebra
    and applied to the n-long planes ONCE: with q_all = q1 T2 orthonormal and the constant in its span, the centred
    columns cc = q_all - one u^T have the Gram matrix I - u u^T, whose range is u-perp, so rest = q_all V with V any
    orthonormal basis of u-perp (a Householder reflector) -- no Gram pass over the samples is needed for it.  9 passes
    over the n x K data instead of ~30.  Ill-conditioned or rank-deficient covariates take the general route.
    """
    n, K = c.shape
    fast = None
    g = c.T @ c
    w, v = np.linalg.eigh(g)
    wmax = w.max() if w.size else 0.0
    if np.isfinite(wmax) and wmax > 0.0 and w.min() > 1e-7 * wmax:
        q1_t = (v / np.sqrt(w)).T @ c.T                       # [K, n]
        w2, v2 = np.linalg.eigh(q1_t @ q1_t.T)                # second pass: removes the O(cond^2 eps) loss of orthogonality
        if w2.min() > 0.5:
            t2 = (v2 / np.sqrt(w2)) @ v2.T                    # q_all = q1 t2 (never formed)
            inv_sqrt_n = 1.0 / np.sqrt(n)
            u = t2.T @ (q1_t.sum(axis=1) * inv_sqrt_n)        # q_all^T one
            resid = inv_sqrt_n - (t2 @ u) @ q1_t              # one - q_all u, explicitly (1 - |u|^2 cannot resolve 1e-9)
            if np.linalg.norm(resid) > 1e-9:
                return np.ascontiguousarray(t2.T @ q1_t), False
            un = u / np.linalg.norm(u)
            e0 = np.zeros(K)
            e0[0] = -1.0 if un[0] > 0 else 1.0                # reflect s e0 <-> un with |un - s e0| >= 1
            hv = un - e0
            h = np.eye(K) - 2.0 * np.outer(hv, hv) / (hv @ hv)
            q_t = np.empty((K, n))
            q_t[0] = inv_sqrt_n
            if K > 1:
                rest = q_t[1:]
                np.matmul((t2 @ h[:, 1:]).T, q1_t, out=rest)  # columns 1.. of the reflector span u-perp
                rest -= rest.mean(axis=1, keepdims=True)      # exact-zero sums up to roundoff
                g3 = rest @ rest.T
                if np.abs(g3 - np.eye(K - 1)).max() > 1e-12:  # (not expected: one more symmetric orthonormalisation)
                    w3, v3 = np.linalg.eigh(g3)
                    if w3.min() > 0.5:
                        rest[:] = ((v3 / np.sqrt(w3)) @ v3.T) @ rest
                        rest -= rest.mean(axis=1, keepdims=True)
                        fast = (q_t, True)
                else:
                    fast = (q_t, True)
            else:
                fast = (q_t, True)
    if fast is not None:
        return fast
    q, has_intercept = orthonormal_basis(c)
    return np.ascontiguousarray(q.T), has_intercept


def _run_device(genotypes, bases, kernel="auto", want_log10_p=False, chunk_variants=None, guard=True):
    """Push the group bases and sweep all rows.  Returns per-group dicts of torch CUDA tensors.

    `guard=False` switches the tolerance guard of the quantised sweeps off for this call (the PCA's power iteration
    only consumes y_transpose_x and corrects itself; include/lrr_b200.h lrr_set_guard)."""
    dev = genotypes.device
    ctx = _lib.context(dev.index)
    M, N = genotypes.n_variants, genotypes.n_samples
    with torch.cuda.device(dev):
        ctx.check(ctx.lib.lrr_set_guard(ctx.handle, 1 if guard else 0))
        _push_groups(ctx, N, bases)
        outs = []
        for b in bases:
            o = {
                "n": torch.empty(M, dtype=torch.int32, device=dev),
                "n_missing": torch.empty(M, dtype=torch.int32, device=dev),
                "sum_x": torch.empty(M, dtype=torch.float64, device=dev),
            }
            for f in STAT_FIELDS:
                o[f] = torch.empty((M, b.P), dtype=torch.float64, device=dev)
            if want_log10_p:
                o["log10_p"] = torch.empty((M, b.P), dtype=torch.float64, device=dev)
            outs.append(o)
        step = M if not chunk_variants else int(chunk_variants)
        ctx.check(ctx.lib.lrr_reserve(ctx.handle, min(M, step) if M else 0))
        stream = torch.cuda.current_stream(dev).cuda_stream
        kid = _lib.KERNELS[kernel]
        for lo in range(0, M, max(step, 1)):
            hi = min(M, lo + step)
            arr = (_lib.GroupOut * len(bases))()
            for g, (b, o) in enumerate(zip(bases, outs)):
                arr[g].n = o["n"][lo:hi].data_ptr()
                arr[g].n_missing = o["n_missing"][lo:hi].data_ptr()
                arr[g].sum_x = o["sum_x"][lo:hi].data_ptr()
                for f in STAT_FIELDS:
                    setattr(arr[g], f, o[f][lo:hi].data_ptr())
                arr[g].log10_p = o["log10_p"][lo:hi].data_ptr() if want_log10_p else None
            ctx.check(ctx.lib.lrr_run(ctx.handle, genotypes.data[lo:hi].data_ptr(), genotypes.flags_ptr(lo), hi - lo,
                                      genotypes.stride, N, arr, len(bases), kid, stream))
        if not guard:
            ctx.check(ctx.lib.lrr_set_guard(ctx.handle, 1))
    return outs


def _add_group(ctx, N, b):
    q = b.q_cols.ctypes.data if b.q_cols.size else None
    qty = b.qty.ctypes.data if b.qty.size else None
    if getattr(b, "weighted", False):
        ctx.check(ctx.lib.lrr_add_group_weighted(ctx.handle, N, b.n, b.K, b.P, b.complete_idx.ctypes.data, q,
                                                 b.y_res.ctypes.data, qty, b.yyp.ctypes.data, b.sqrt_w.ctypes.data))
    else:
        ctx.check(ctx.lib.lrr_add_group(ctx.handle, N, b.n, b.K, b.P, int(b.has_intercept), b.complete_idx.ctypes.data,
                                        q, b.y_res.ctypes.data, qty, b.yyp.ctypes.data))


def _push_groups(ctx, N, bases):
    ctx.check(ctx.lib.lrr_clear_groups(ctx.handle))
    for b in bases:
        _add_group(ctx, N, b)


class _HostStream:
    """lrr_stream_* over a HostBedGenotypes: begun before the driver prologue so that the first blocks cross PCIe
    while the host computes the covariate basis."""

    def __init__(self, genotypes, block_variants=0, depth=0):
        self.g = genotypes
        self.ctx = _lib.context(genotypes.device.index)
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(genotypes.device):
            # groups of an earlier call are dropped first: lrr_clear_groups synchronises the device, which must not
            # happen behind the copies queued by lrr_stream_begin
            self.ctx.check(self.ctx.lib.lrr_clear_groups(self.ctx.handle))
            self.ctx.check(self.ctx.lib.lrr_stream_begin(
                self.ctx.handle, ctypes.byref(self.handle), genotypes.rows.data_ptr(), genotypes.n_variants,
                genotypes.bed_stride, genotypes.n_samples, int(block_variants), int(depth)))

    def run(self, bases, kernel="auto", want_log10_p=False):
        """Returns per-group dicts of numpy arrays (views of one page-locked result buffer)."""
        ctx, M, N = self.ctx, self.g.n_variants, self.g.n_samples
        with torch.cuda.device(self.g.device):
            for b in bases:
                _add_group(ctx, N, b)
            fields = [("n", np.int32, 0), ("n_missing", np.int32, 0), ("sum_x", np.float64, 0)]
            fields += [(f, np.float64, 1) for f in STAT_FIELDS + (["log10_p"] if want_log10_p else [])]
            total, layout = 0, []
            for b in bases:
                lay = {}
                for name, dt, per_p in fields:
                    nbytes = M * (b.P if per_p else 1) * np.dtype(dt).itemsize
                    lay[name] = (total, nbytes)
                    total += (nbytes + 255) // 256 * 256
                layout.append(lay)
            buf = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=True)
            base, host = buf.data_ptr(), buf.numpy()
            arr = (_lib.GroupOut * len(bases))()
            outs = []
            for g, (b, lay) in enumerate(zip(bases, layout)):
                o = {}
                for name, dt, per_p in fields:
                    off, nbytes = lay[name]
                    setattr(arr[g], name, base + off)
                    v = host[off:off + nbytes].view(dt)
                    o[name] = v.reshape(M, b.P) if per_p else v
                if not want_log10_p:
                    arr[g].log10_p = None
                outs.append(o)
            ctx.check(ctx.lib.lrr_stream_run(ctx.handle, self.handle, arr, len(bases), _lib.KERNELS[kernel]))
        return outs

    def close(self):
        if self.handle:
            self.ctx.lib.lrr_stream_end(self.ctx.handle, self.handle)
            self.handle = ctypes.c_void_p()


def _run_device_dense(dosage, bases):
    """lrr_run_dense over a DenseDosage (lrr_run_dense_u16 over a CompactDosage).  Returns per-group dicts of torch CUDA tensors."""
    dev = dosage.device
    ctx = _lib.context(dev.index)
    M, N = dosage.n_variants, dosage.n_samples
    with torch.cuda.device(dev):
        _push_groups(ctx, N, bases)
        outs = []
        arr = (_lib.GroupOut * len(bases))()
        for g, b in enumerate(bases):
            o = {"n": torch.empty(M, dtype=torch.int32, device=dev), "n_missing": torch.empty(M, dtype=torch.int32, device=dev),
                 "sum_x": torch.empty(M, dtype=torch.float64, device=dev)}
            for f in STAT_FIELDS:
                o[f] = torch.empty((M, b.P), dtype=torch.float64, device=dev)
            for k, v in o.items():
                setattr(arr[g], k, v.data_ptr())
            arr[g].log10_p = None
            outs.append(o)
        if hasattr(dosage, "scale"):   # CompactDosage: uint16 entries
            ctx.check(ctx.lib.lrr_run_dense_u16(ctx.handle, dosage.data.data_ptr(), M, dosage.ld, N, dosage.scale, arr, len(bases),
                                                torch.cuda.current_stream(dev).cuda_stream))
        else:
            ctx.check(ctx.lib.lrr_run_dense(ctx.handle, dosage.data.data_ptr(), M, N, N, arr, len(bases),
                                            torch.cuda.current_stream(dev).cuda_stream))
    return outs


def linear_regression_rows(y, x, covariates, block_size=16, pass_through=(), *, weights=None,
                           _kernel="auto", _log10_p=False, _stream_block=0, _stream_depth=0, _guard=True,
                           _sharded=False) -> Table:
    """For each row, test an input variable for association with response variables using linear regression.

    Drop-in for `hl.linear_regression_rows` (statgen.py:235): same arguments, same validation, same output
    fields in the same order -- row key, pass_through, then `n, sum_x, y_transpose_x, beta, standard_error,
    t_stat, p_value` (scalars when `y` is one expression, arrays of length P for a list, arrays over groups for
    a list of lists).  `block_size` is accepted and numerically inert on the GPU.  `weights` (one expression, or one
    per group of a chained `y`): weighted least squares as the reference's `_linear_regression_rows_nd` does it
    (statgen.py:557-581, 636-660): samples without a weight are dropped, x is mean-imputed and then scaled by sqrt(w).

    Like the reference (statgen.py:370-401) the call selects its inputs into uniquely named fields, builds the config
    `{'name': 'LinearRegressionRowsSingle' | '...Chained', 'yFields', 'xField', 'covFields', 'rowBlockSize',
    'passThrough'}` and applies the relational function registered under that name (hail_b200/plugin.py, the
    analogue of `MatrixToTableApply` + RelationalFunctions.scala:112-138).

    `_sharded=True` (one process per GPU, torch.distributed initialised): every rank passes a MatrixTable holding its own
    contiguous range of the ROWS (variants) over the same columns -- the reference's one task per partition (LR:95,
    :274).  Rank 0 runs the driver prologue and its bases are broadcast once per call (LR:74-78, :257); every rank
    sweeps its rows and returns the Table of ITS rows (bit-identical to the same rows of a single-device run);
    `Table.gather()` concatenates all ranks' rows in rank order on every rank.  `_sharded="peer"` sends the broadcast
    through symmetric memory (pulled over NVLink by the copy engines, hail_b200/dist.py PeerBroadcast) instead of NCCL.
    """
    if not isinstance(block_size, int):
        raise TypeError("linear_regression_rows: 'block_size' must be int")
    if not isinstance(x, EntryExpression):
        raise ExpressionException("'linear_regression_rows/x': expected an entry-indexed expression "
                                  "(e.g. mt.GT.n_alt_alleles())")
    mt = x.source

    y_is_list = isinstance(y, (list, tuple))
    if y_is_list and len(y) == 0:
        raise ValueError("'linear_regression_rows': found no values for 'y'")  # SG:351-352
    is_chained = y_is_list and isinstance(y[0], (list, tuple))
    if is_chained and any(len(lst) == 0 for lst in y):
        raise ValueError("'linear_regression_rows': found empty inner list for 'y'")  # SG:354-355

    if weights is not None:   # statgen.py:437-467
        if y_is_list and is_chained and not isinstance(weights, list):
            raise ValueError("When y is a list of lists, weights should be a list.")
        elif y_is_list and not is_chained and isinstance(weights, list):
            raise ValueError("When y is a single list, weights should be a single expression.")
        elif not y_is_list and isinstance(weights, list):
            raise ValueError("When y is a single expression, weights should be a single expression.")
        weights = weights if isinstance(weights, list) else [weights]
        if len(weights) != (len(y) if is_chained else 1):
            raise ValueError("Must specify same number of weights as groups of phenotypes")

    groups = [list(g) for g in y] if is_chained else [list(y) if y_is_list else [y]]
    y_vals = [[_column_values(e, mt, "linear_regression_rows/y") for e in g] for g in groups]
    cov_vals = [_column_values(e, mt, "linear_regression_rows/covariates") for e in covariates]
    w_vals = None if weights is None else [_column_values(e, mt, "linear_regression_rows/weights") for e in weights]
    _warn_if_no_intercept("linear_regression_rows", covariates)
    row_fields = _get_regression_row_fields(mt, pass_through, "linear_regression_rows")

    # SG:370-392: select the inputs into uniquely named column / entry fields ...
    x_field_name = f"__uid_x_{next(_uid)}"
    if is_chained:
        y_field_names = [[f"__y_{i}_{j}" for j in range(len(g))] for i, g in enumerate(y_vals)]
    else:
        y_field_names = [f"__y_{i}" for i in range(len(y_vals[0]))]
    cov_field_names = [f"__cov{i}" for i in range(len(cov_vals))]
    flat_names = list(itertools.chain.from_iterable(y_field_names)) if is_chained else y_field_names
    selected = mt._copy(cols=OrderedDict(itertools.chain(zip(flat_names, itertools.chain.from_iterable(y_vals)),
                                                         zip(cov_field_names, cov_vals))),
                        rows=OrderedDict(itertools.chain(((k, mt.row[k]) for k in mt.row_key), row_fields.items())),
                        col_key=()).annotate_entries(**{x_field_name: x})
    # ... SG:394-401: and hand the config to the relational function registered under its name
    config = {
        "name": "LinearRegressionRowsChained" if is_chained else "LinearRegressionRowsSingle",
        "yFields": y_field_names,
        "xField": x_field_name,
        "covFields": cov_field_names,
        "rowBlockSize": block_size,
        "passThrough": [f for f in row_fields if f not in mt.row_key],
    }
    from . import plugin
    ht = plugin.matrix_to_table_apply(selected, config, weights=w_vals, kernel=_kernel, log10_p=_log10_p,
                                      stream_block=_stream_block, stream_depth=_stream_depth, guard=_guard, sharded=_sharded)
    if not y_is_list:   # SG:404-406
        fields = STAT_FIELDS + (["log10_p"] if _log10_p else [])
        ht = ht.annotate(**{f: ht[f][:, 0] for f in fields})
    return ht


_uid = itertools.count()
LAST_STREAM_PHASES = {}   # host-side phase times of the last call that streamed host-resident rows (measurement aid)


def _execute(mt, x, y_vals, cov_vals, is_chained, pass_through_names, *, weights=None, kernel="auto", log10_p=False,
             stream_block=0, stream_depth=0, guard=True, sharded=False) -> Table:
    """The body of LinearRegressionRowsSingle / Chained `execute` (LR:46-195 / 226-407): driver prologue on the host,
    the per-partition loop on the device.  `y_vals` is a list of groups, each a list of float64 column arrays; returns
    the Table with array-valued statistics ([M, P] per group; a ChainedField over groups when `is_chained`)."""
    n_cols = mt.count_cols()
    w_vals = [None] * len(y_vals) if weights is None else weights
    from .genotypes import CompactDosage, DenseDosage, HostBedGenotypes
    DenseDosage = (DenseDosage, CompactDosage)   # both are dense entry fields; the device call differs (_run_device_dense)

    def make_bases():
        # (built here, not above: a streamed call has its first host -> device copies in flight by now.)  The covariates are
        # stacked as contiguous planes [K, n] and handed over as the transposed VIEW: the prologue works plane by plane
        cov = np.stack(cov_vals).T if cov_vals else np.empty((n_cols, 0))
        return [GroupBasis(np.column_stack(g), cov, mt.col_index, i if is_chained else None, w_vals[i])
                for i, g in enumerate(y_vals)]

    if isinstance(mt.genotypes, DenseDosage) or x.kind == "dosage":
        if not isinstance(mt.genotypes, DenseDosage) or x.kind != "dosage":
            raise ExpressionException("'linear_regression_rows/x': a dense dosage field needs a DenseDosage entry matrix")
        outs = _run_device_dense(mt.genotypes, make_bases())
        torch.cuda.synchronize(mt.genotypes.device)
        host = [{k: v.cpu().numpy() for k, v in o.items()} for o in outs]
    elif isinstance(mt.genotypes, HostBedGenotypes) and mt.genotypes.nbytes >= _STREAM_MIN_BYTES:
        # host-resident .bed rows: stream them through the device; the copies start before the prologue
        t0 = time.perf_counter()
        stream = _HostStream(mt.genotypes, stream_block, stream_depth)
        try:
            t1 = time.perf_counter()
            bases = make_bases()
            t2 = time.perf_counter()
            host = stream.run(bases, kernel=kernel, want_log10_p=log10_p)
            t3 = time.perf_counter()
        finally:
            stream.close()
        # where the wall time of the last streamed call went (host side; the copies run underneath all of it)
        LAST_STREAM_PHASES.update(stream_begin_ms=1e3 * (t1 - t0), host_prologue_ms=1e3 * (t2 - t1),
                                  add_groups_and_stream_run_ms=1e3 * (t3 - t2), close_ms=1e3 * (time.perf_counter() - t3))
    elif sharded:
        import torch.distributed as tdist

        from .dist import ShardedRegression
        if weights is not None:
            raise NotImplementedError("linear_regression_rows: weights= with _sharded=True")
        sr = ShardedRegression(mt.genotypes, transport="peer" if sharded == "peer" else "nccl")
        sr.set_bases(make_bases() if tdist.get_rank() == 0 else None)
        outs = sr.run(kernel=kernel, want_log10_p=log10_p)
        torch.cuda.synchronize(mt.genotypes.device)
        host = [{k: v.cpu().numpy() for k, v in o.items()} for o in outs]
    else:
        # (a small host-resident .bed goes to the device in one piece: the streaming machinery -- four streams, a slot
        # ring, a page-locked result buffer -- costs more than a sweep of a few megabytes)
        g = mt.genotypes.to_device() if isinstance(mt.genotypes, HostBedGenotypes) else mt.genotypes
        outs = _run_device(g, make_bases(), kernel=kernel, want_log10_p=log10_p, guard=guard)
        torch.cuda.synchronize(g.device)
        host = [{k: v.cpu().numpy() for k, v in o.items()} for o in outs]

    fields = OrderedDict()
    for k in mt.row_key:
        fields[k] = mt.row[k]
    for k in pass_through_names:
        fields[k] = mt.row[k]
    stat_names = STAT_FIELDS + (["log10_p"] if log10_p else [])
    if is_chained:  # LR:208-216
        fields["n"] = np.stack([h["n"] for h in host], axis=1)
        fields["sum_x"] = np.stack([h["sum_x"] for h in host], axis=1)
        for f in stat_names:
            fields[f] = ChainedField(h[f] for h in host)
    else:           # LR:26-34: array<float64> of length P per row
        h = host[0]
        fields["n"] = h["n"]
        fields["sum_x"] = h["sum_x"]
        for f in stat_names:
            fields[f] = h[f]
    t = Table(fields, key=mt.row_key, n_rows=mt.count_rows())
    t.n_missing = [h["n_missing"] for h in host] if is_chained else host[0]["n_missing"]
    t.sharded = bool(sharded)
    return t


_STREAM_MIN_BYTES = 64 << 20


def lambda_gc(p_value, approximate=True, device=0) -> float:
    """Genomic inflation factor of a set of p-values: median(qchisqtail(p, 1)) / qchisqtail(0.5, 1) over the non-NaN
    p-values (drop-in for `hl.lambda_gc`, statgen.py:3096-3128; `approximate` is accepted and ignored -- the median is
    exact here).  The quantile function runs on the device (lrr_qchisqtail1), the order statistic in torch."""
    p = np.ascontiguousarray(np.asarray(p_value, dtype=np.float64).reshape(-1))
    dev = torch.device("cuda", device)
    ctx = _lib.context(dev.index)
    with torch.cuda.device(dev):
        d_p = torch.from_numpy(p).to(dev)
        d_p = d_p[~torch.isnan(d_p)].contiguous()
        if d_p.numel() == 0:
            return float("nan")
        chi = torch.empty_like(d_p)
        ctx.check(ctx.lib.lrr_qchisqtail1(ctx.handle, d_p.data_ptr(), d_p.numel(), chi.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream))
        srt = torch.sort(chi).values
        n = srt.numel()
        med = srt[n // 2] if n % 2 else 0.5 * (srt[n // 2 - 1] + srt[n // 2])
        return float(med) / 0.454936423119572   # qchisqtail(0.5, 1)
